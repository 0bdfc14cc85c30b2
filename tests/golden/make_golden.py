"""Generates tests/golden/vtrace_golden.npz.  Run ONLY in the build container
(where /root/reference exists):   python tests/golden/make_golden.py

What it executes (nothing is copied into this repo):
  * the UNMODIFIED reference source common/vtrace.py and
    agents/policy_gradient/modules/advantages.py, imported from
    /root/reference over tf_numpy_shim (numpy fp32 stand-ins for the TF ops);
  * the reference test's own numpy ground truth
    tests/vtrace_test.py:41-82 `_ground_truth_calculation`, pulled out of the
    test file by AST (the file's top-level imports need TensorFlow) and run as is.
The outputs are stored next to the inputs so that the GPU box (which has no
/root/reference) can replay them.
"""
import ast
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402


def _load(path, name):
  spec = importlib.util.spec_from_file_location(name, path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


def _extract_function(path, fn_name, namespace):
  tree = ast.parse(open(path).read())
  for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name == fn_name:
      code = compile(ast.Module(body=[node], type_ignores=[]), path, 'exec')
      exec(code, namespace)
      return namespace[fn_name]
  raise KeyError(fn_name)


def main():
  tf_numpy_shim.install()
  ref_vtrace = _load(os.path.join(REF, 'common/vtrace.py'), 'ref_vtrace')
  ref_adv = _load(os.path.join(
      REF, 'agents/policy_gradient/modules/advantages.py'), 'ref_advantages')
  ns = {'np': np, 'vtrace': ref_vtrace}
  ground_truth = _extract_function(
      os.path.join(REF, 'tests/vtrace_test.py'), '_ground_truth_calculation', ns)

  out = {}
  f32 = np.float32

  # ---- case A: known-answer case of tests/vtrace_test.py:118-145 ----------
  T = B = 5
  ar = np.arange(T * B, dtype=f32).reshape(T, B)
  log_rhos = f32(5) * (ar / f32(B * T) - f32(0.5))
  a = dict(
      behaviour_action_log_probs=np.zeros_like(log_rhos),
      target_action_log_probs=log_rhos,
      discounts=np.array([[0.9 / (b + 1) for b in range(B)] for _ in range(T)]),
      rewards=ar.copy(),
      values=ar / f32(B),
      bootstrap_value=np.arange(B, dtype=f32) + f32(1.0),
      clip_rho_threshold=3.7,
      clip_pg_rho_threshold=2.2)
  r = ref_vtrace.from_importance_weights(**a)
  g = ground_truth(**a)     # float64 discounts, as in the reference test
  for k, v in a.items():
    out['A_' + k] = np.asarray(v)
  out['A_ref_vs'] = np.asarray(r.vs.a, f32)
  out['A_ref_pg'] = np.asarray(r.pg_advantages.a, f32)
  out['A_gt_vs'] = np.asarray(g.vs)
  out['A_gt_pg'] = np.asarray(g.pg_advantages)

  # ---- case B: BASELINE cfg 1: T=20 B=64 seed 0 ----------------
  rng = np.random.default_rng(0)
  T, B = 20, 64
  b = dict(
      target_action_log_probs=rng.uniform(-2, 2, (T, B)).astype(f32),
      behaviour_action_log_probs=rng.uniform(-2, 2, (T, B)).astype(f32),
      discounts=(0.99 * (rng.random((T, B)) < 0.95)).astype(f32),
      rewards=rng.uniform(0, 3, (T, B)).astype(f32),
      values=rng.uniform(0, 3, (T, B)).astype(f32),
      bootstrap_value=rng.uniform(0, 3, (B,)).astype(f32))
  for name, kw in (('B1', {}),                                # default clips 1/1
                   ('B2', dict(clip_rho_threshold=None,
                               clip_pg_rho_threshold=None)),   # no clipping
                   ('B3', dict(clip_rho_threshold=3.7,
                               clip_pg_rho_threshold=2.2, lambda_=0.9))):
    r = ref_vtrace.from_importance_weights(**b, **kw)
    out[name + '_ref_vs'] = np.asarray(r.vs.a, f32)
    out[name + '_ref_pg'] = np.asarray(r.pg_advantages.a, f32)
  for k, v in b.items():
    out['B_' + k] = v

  # ---- case C: advantages_test.py:129-150 (lambda=0.95, done-masked) -------
  rng = np.random.default_rng(1)
  values = rng.uniform(0, 3, (21, 10)).astype(f32)
  rewards = rng.uniform(0, 3, (20, 10)).astype(f32)
  tlp = rng.uniform(-2, 2, (20, 10)).astype(f32)
  blp = rng.uniform(-2, 2, (20, 10)).astype(f32)
  done = rng.random((20, 10)) < 0.05
  T_ = tf_numpy_shim.Tensor
  targets, _ = ref_adv.vtrace(T_(values), T_(rewards), T_(done),
                              T_(np.zeros_like(done)), 0.99, T_(tlp), T_(blp),
                              lambda_=0.95)
  seed = ref_vtrace.from_importance_weights(
      T_(tlp), T_(blp), T_((0.99 * (~done)).astype(f32)), T_(rewards),
      T_(values[:-1]), T_(values[-1]), lambda_=0.95)
  out.update(C_values=values, C_rewards=rewards, C_tlp=tlp, C_blp=blp,
             C_done=done, C_adv_targets=np.asarray(targets.a, f32),
             C_ref_vs=np.asarray(seed.vs.a, f32),
             C_ref_pg=np.asarray(seed.pg_advantages.a, f32))

  # ---- case D: extra trailing dim [T,B,C] (vtrace.py:49-51) ----------------
  rng = np.random.default_rng(2)
  T, B, C = 7, 3, 2
  d = dict(
      target_action_log_probs=rng.uniform(-1, 1, (T, B, C)).astype(f32),
      behaviour_action_log_probs=rng.uniform(-1, 1, (T, B, C)).astype(f32),
      discounts=np.full((T, B, C), 0.97, f32),
      rewards=rng.normal(size=(T, B, C)).astype(f32),
      values=rng.normal(size=(T, B, C)).astype(f32),
      bootstrap_value=rng.normal(size=(B, C)).astype(f32))
  r = ref_vtrace.from_importance_weights(**d)
  for k, v in d.items():
    out['D_' + k] = v
  out['D_ref_vs'] = np.asarray(r.vs.a, f32)
  out['D_ref_pg'] = np.asarray(r.pg_advantages.a, f32)

  path = os.path.join(HERE, 'vtrace_golden.npz')
  np.savez_compressed(path, **out)
  print('wrote', path, {k: v.shape for k, v in out.items() if hasattr(v, 'shape')})


if __name__ == '__main__':
  main()
