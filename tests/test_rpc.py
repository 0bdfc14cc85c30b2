"""CPU: the RPC layer (seed_rl_b200/grpc/ops.py) -- wire codec, Server/Client over a real
gRPC channel with the C++ batcher underneath (unpinned slabs, numpy compute function), and
the reference's batching / error behaviours (grpc/python/ops_test.py)."""
import threading

import numpy as np
import pytest

from seed_rl_b200.common import utils
from seed_rl_b200.grpc import ops

TS = utils.TensorSpec


def test_tensor_proto_roundtrip_and_known_bytes():
  a = np.arange(6, dtype=np.int32).reshape(2, 3)
  buf = ops.encode_tensor(a)
  # dtype=DT_INT32(3); shape {dim{size:2} dim{size:3}}; tensor_content = 24 raw bytes
  assert buf[:2] == b'\x08\x03' and buf[2:4] == b'\x12\x08'
  assert buf[4:12] == b'\x12\x02\x08\x02\x12\x02\x08\x03'
  assert buf[12:14] == b'\x22\x18' and buf[14:] == a.tobytes()
  for x in (a, np.float32(3.5), np.zeros((0, 4), np.uint8), np.array([True, False]),
            np.arange(5, dtype=np.int64) - 2, np.random.default_rng(0).normal(size=(3, 2, 2))):
    y = ops.decode_tensor(ops.encode_tensor(np.asarray(x)))
    assert y.dtype == np.asarray(x).dtype and y.shape == np.asarray(x).shape
    np.testing.assert_array_equal(y, x)


def test_structured_value_roundtrip():
  specs = (TS([], 'int64', 'action'), [TS([18], 'float32', 'logits'), TS([2, 3], 'uint8', None)])
  out = ops.decode_structure(ops.encode_structure(specs))
  assert isinstance(out, tuple) and isinstance(out[1], list)
  assert out[0] == TS([], 'int64', 'action') and out[1][0] == TS([18], 'float32', 'logits')
  assert out[1][1].shape == [2, 3] and out[1][1].dtype == 'uint8'
  assert ops.decode_structure(ops.encode_structure(None)) is None


@pytest.fixture
def server_client(tmp_path):
  made = []

  def make(fn, address=None):
    address = address or 'unix:%s' % (tmp_path / ('sock%d' % len(made)))
    server = ops.Server([address], pinned=False)
    server.bind(fn)
    server.start()
    made.append(server)
    return server, address
  yield make
  for s in made:
    s.shutdown()


def test_batching_single_calls_fill_a_batch(server_client):
  """reference ops_test.py: N single-element calls are batched into one [N] invocation."""
  seen = []

  @ops.function((TS([4], 'int32', 'x'),), TS([4], 'int32', 'y'))
  def foo(x):
    seen.append(np.array(x))
    return x + 1
  server, address = server_client(foo)
  clients = [ops.Client(address) for _ in range(4)]
  out = [None] * 4

  def call(i):
    out[i] = clients[i].foo(np.int32(10 * i))
  ts = [threading.Thread(target=call, args=(i,)) for i in range(4)]
  [t.start() for t in ts]; [t.join(20) for t in ts]
  assert [int(o) for o in out] == [1, 11, 21, 31]
  assert len(seen) == 1 and sorted(seen[0].tolist()) == [0, 10, 20, 30]
  assert all(o.shape == () for o in out)            # unbatched call -> unbatched reply


def test_prebatched_slices_and_nest_arguments(server_client):
  """reference ops_test.py:776-799 ([2]+[2] -> [4]) and :356-382 (nests)."""
  @ops.function((TS([4], 'int32', 'a'), (TS([4, 3], 'float32', 'b'), TS([4], 'bool', 'c'))),
                (TS([4], 'int32', 'o0'), TS([4, 3], 'float32', 'o1')))
  def foo(a, bc):
    b, c = bc
    return a * 2, b + c[:, None]
  server, address = server_client(foo)
  c1, c2 = ops.Client(address), ops.Client(address)
  res = {}

  def call(cl, key, a):
    b = np.full((2, 3), a[0], np.float32)
    res[key] = cl.foo(np.asarray(a, np.int32), (b, np.array([True, False])))
  ts = [threading.Thread(target=call, args=(c1, 1, [1, 2])), threading.Thread(target=call, args=(c2, 2, [3, 4]))]
  [t.start() for t in ts]; [t.join(20) for t in ts]
  assert res[1][0].tolist() == [2, 4] and res[2][0].tolist() == [6, 8]
  np.testing.assert_array_equal(res[1][1], [[2, 2, 2], [1, 1, 1]])
  np.testing.assert_array_equal(res[2][1], [[4, 4, 4], [3, 3, 3]])


def test_error_strings(server_client):
  """reference grpc/ops/grpc.cc:513-549,187-190 via ops_test.py:303-336,564-630."""
  @ops.function((TS([2, 3], 'int32', 'x'),), TS([2], 'int32', 'y'))
  def foo(x):
    return x.sum(-1)
  server, address = server_client(foo)
  c = ops.Client(address)
  with pytest.raises(ops.InvalidArgumentError, match='Expects 1 arguments, but 2 is provided'):
    c.foo(np.zeros(3, np.int32), np.zeros(3, np.int32))
  with pytest.raises(ops.InvalidArgumentError, match=r'Expects arg\[0\] to be int32 but float is provided'):
    c.foo(np.zeros(3, np.float32))
  with pytest.raises(ops.InvalidArgumentError, match=r'Expects arg\[0\] to have shape with suffix \[3\], but had shape \[4\]'):
    c.foo(np.zeros(4, np.int32))
  with pytest.raises(ops.InvalidArgumentError, match=r'to have shape with 1 dimension\(s\), but had shape \[1, 1, 3\]'):
    c.foo(np.zeros((1, 1, 3), np.int32))
  with pytest.raises(ops.RpcError, match='Function bar not found'):
    server.call_local('bar', [])
  # exact [N, ...] input = one whole batch from one caller (grpc.cc:626-629)
  assert c.foo(np.arange(6, dtype=np.int32).reshape(2, 3)).tolist() == [3, 12]


def test_shutdown_cancels_partial_batch_and_client_sees_unavailable(server_client):
  """reference ops_test.py:384-501: a waiter on a partially filled batch is released by
  shutdown and the client raises UnavailableError."""
  @ops.function((TS([2], 'int32', 'x'),), TS([2], 'int32', 'y'))
  def foo(x):
    return x
  server, address = server_client(foo)
  c = ops.Client(address)
  err = []

  def call():
    try:
      c.foo(np.int32(1))      # never completes: the batch of 2 is never filled
    except ops.UnavailableError as e:
      err.append(e)
  t = threading.Thread(target=call); t.start()
  import time; time.sleep(0.3)
  server.shutdown()
  t.join(20)
  assert err and 'server closed' in str(err[0])


def test_bind_requires_signature_and_batched_outputs():
  server = ops.Server(['unix:/tmp/seedrl_b200_unused'], pinned=False)
  with pytest.raises(ValueError, match='input_signature'):
    server.bind(lambda x: x)

  @ops.function((TS([4], 'int32', 'x'),), TS([3], 'int32', 'y'))
  def bad(x):
    return x
  with pytest.raises(ValueError, match='first dimension 4'):
    server.bind(bad)


def test_service_envelope_matches_the_reference_descriptor():
  """Every message of grpc/service.proto, serialised by the protobuf runtime from the
  reference's own compiled descriptor (tests/golden/make_golden_rpc.py), against the
  hand-written envelope codec -- byte-identical encodings and lossless decodings."""
  import json, os
  from seed_rl_b200.grpc import ops
  g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'rpc_golden.json')))
  # the schema itself: field numbers / wire types the codec hard-codes
  f = {m: {name: (num, typ, lab) for name, num, typ, lab in fs} for m, fs in g['descriptor_fields'].items()}
  assert f['CallRequest'] == {'function': (1, 9, 1), 'tensor': (2, 12, 3)}
  assert f['CallResponse'] == {'tensor': (1, 12, 3), 'status_code': (2, 5, 1), 'status_error_message': (3, 9, 1)}
  assert f['MethodOutputSignature'] == {'name': (1, 9, 1), 'output_specs': (2, 12, 1)}
  assert f['InitResponse'] == {'method_output_signature': (1, 11, 3)}
  seen = set()
  for c in g['cases']:
    want = bytes.fromhex(c['hex']); fl = c['fields']; seen.add(c['kind'])
    if c['kind'] == 'CallRequest':
      tensors = [bytes.fromhex(t) for t in fl['tensor']]
      assert ops._encode_call_request_raw(fl['function'], tensors) == want
      assert ops._decode_call_request_raw(want) == (fl['function'], tensors)
    elif c['kind'] == 'CallResponse':
      tensors = [bytes.fromhex(t) for t in fl['tensor']]
      assert ops._encode_call_response_raw(tensors, fl['status_code'], fl['status_error_message']) == want
      assert ops._decode_call_response_raw(want) == (tensors, fl['status_code'], fl['status_error_message'])
    elif c['kind'] == 'InitResponse':
      sigs = list(zip(fl['names'], [bytes.fromhex(s) for s in fl['specs']]))
      assert ops._encode_init_response_raw(sigs) == want
      assert ops._decode_init_response_raw(want) == sigs
    elif c['kind'] == 'InitRequest':
      assert want == b''                   # what Client sends to /Init
  assert seen == {'CallRequest', 'CallResponse', 'InitResponse', 'InitRequest'}


def test_tensorproto_and_structuredvalue_match_tf_schemas():
  """encode_tensor / encode_structure against bytes produced by the protobuf runtime from
  protoc-compiled copies of TensorFlow's tensor.proto / struct.proto
  (tests/golden/make_golden_tfproto.py), both directions."""
  import json, os
  from seed_rl_b200.common.utils import TensorSpec
  from seed_rl_b200.grpc import ops
  g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'tfproto_golden.json')))
  assert g['enum'] == {'float32': 1, 'int32': 3, 'uint8': 4, 'int64': 9, 'bool': 10}
  for t in g['tensors']:
    a = np.frombuffer(bytes.fromhex(t['content']), dtype=t['dtype']).reshape(t['shape'])
    want = bytes.fromhex(t['hex'])
    got = ops.encode_tensor(a)
    back = ops.decode_tensor(want)
    assert back.dtype == a.dtype and back.shape == a.shape and np.array_equal(back, a)
    assert got == want, (t['dtype'], t['shape'])          # byte-identical, proto3 default omission included

  def build(j):
    if j is None:
      return None
    if 'tuple' in j:
      return tuple(build(x) for x in j['tuple'])
    if 'list' in j:
      return [build(x) for x in j['list']]
    return TensorSpec(j['shape'], j['dtype'], j['name'])

  def same(a, b):
    if a is None or b is None:
      return a is None and b is None
    if isinstance(a, TensorSpec):
      return (isinstance(b, TensorSpec) and list(a.shape) == list(b.shape) and np.dtype(a.dtype) == np.dtype(b.dtype)
              and (a.name or None) == (b.name or None))
    return type(a) is type(b) and len(a) == len(b) and all(same(x, y) for x, y in zip(a, b))
  for s in g['structures']:
    spec = build(s['spec'])
    want = bytes.fromhex(s['hex'])
    assert same(ops.decode_structure(want), spec), s['spec']
    assert ops.encode_structure(spec) == want, s['spec']    # byte-identical


def test_decoders_reject_malformed_bytes_cleanly():
  """Random bytes, truncations and bit flips of valid messages either decode or raise
  ops.MalformedProto (a ValueError) -- never IndexError / MemoryError / a hang; a malformed
  CallRequest on the wire comes back as InvalidArgument, like a failed ParseFromString."""
  from seed_rl_b200.grpc import ops
  rng = np.random.default_rng(0)
  decoders = (ops.decode_tensor, ops.decode_structure, ops._decode_call_request_raw,
              ops._decode_call_response_raw, ops._decode_init_response_raw, ops._decode_call_request)
  valid = [ops.encode_tensor(rng.normal(size=(3, 4)).astype(np.float32)),
           ops.encode_structure((TS([2, 3], 'float32', 'x'), [TS([], 'int64', None)])),
           ops._encode_call_request('inference', [np.arange(5, dtype=np.int32), np.zeros((2, 2), np.uint8)]),
           ops._encode_call_response([np.ones(3, np.float32)], 3, 'bad'),
           ops._encode_init_response_raw([('inference', bytes([0x9a, 0x02, 0x00]))])]
  cases = [bytes(rng.integers(0, 256, int(rng.integers(0, 48)), dtype=np.uint8)) for _ in range(4000)]
  for v in valid:
    for _ in range(300):
      b = bytearray(v)
      op = rng.integers(0, 3)
      if op == 0 and len(b) > 1:
        b = b[:int(rng.integers(0, len(b)))]
      elif op == 1:
        b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
      else:
        i = int(rng.integers(0, len(b) + 1)); b[i:i] = bytes(rng.integers(0, 256, 3, dtype=np.uint8))
      cases.append(bytes(b))
  outcomes = set()
  for c in cases:
    for d in decoders:
      try:
        d(c); outcomes.add('ok')
      except ops.MalformedProto:
        outcomes.add('malformed')
  assert outcomes == {'ok', 'malformed'}


def test_malformed_call_request_is_invalid_argument_on_the_stream(server_client):
  from seed_rl_b200.grpc import ops
  @ops.function((TS([4], 'int32', 'x'),), TS([4], 'int32', 'y'))
  def foo(x):
    return x + 1
  server, _ = server_client(foo)
  resp = list(server._call_rpc(iter([bytes([0x0a, 0xff, 0xff, 0xff, 0xff, 0x0f])]), None))     # length runs past the end
  tensors, code, msg = ops._decode_call_response_raw(resp[0])
  assert tensors == [] and code == ops.INVALID_ARGUMENT and msg.startswith('Malformed CallRequest')
