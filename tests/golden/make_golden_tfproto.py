"""Generates tests/golden/tfproto_golden.json.  Run in the build container:
    python tests/golden/make_golden_tfproto.py

The two TensorFlow payload formats carried inside seed_rl.TensorService messages --
tensorflow.TensorProto (as written by Tensor::AsProtoTensorContent, grpc/ops/grpc.cc:160-184)
and tensorflow.StructuredValue / TensorSpecProto (output_specs, grpc.cc:1145-1179) -- are not
vendored under /root/reference and TensorFlow is not installable here.  TensorBoard 2.20 ships
protoc-compiled copies of exactly these TF schemas (tensorboard/compat/proto/{tensor,struct,
tensor_shape,types}_pb2.py: same field numbers, package renamed); this script serialises
sample messages with the protobuf runtime from those schemas.  tests/test_rpc.py then checks
seed_rl_b200.grpc.ops.encode_tensor / encode_structure against the bytes in both directions,
so the test box needs neither TensorFlow nor TensorBoard."""
import json
import os

import numpy as np
from tensorboard.compat.proto import struct_pb2, tensor_pb2, tensor_shape_pb2, types_pb2

HERE = os.path.dirname(os.path.abspath(__file__))
DT = {'float32': types_pb2.DT_FLOAT, 'int32': types_pb2.DT_INT32, 'uint8': types_pb2.DT_UINT8,
      'int64': types_pb2.DT_INT64, 'bool': types_pb2.DT_BOOL}


def shape_proto(shape):
  return tensor_shape_pb2.TensorShapeProto(dim=[tensor_shape_pb2.TensorShapeProto.Dim(size=int(d)) for d in shape])


def tensor_proto(a):
  return tensor_pb2.TensorProto(dtype=DT[a.dtype.name], tensor_shape=shape_proto(a.shape),
                                tensor_content=a.tobytes())


def spec_value(spec):
  """nest of (name, shape, dtype) | None | list | tuple -> StructuredValue."""
  v = struct_pb2.StructuredValue()
  if spec is None:
    v.none_value.SetInParent()
  elif isinstance(spec, dict):
    v.tensor_spec_value.name = spec['name'] or ''
    v.tensor_spec_value.shape.CopyFrom(shape_proto(spec['shape']))
    v.tensor_spec_value.dtype = DT[spec['dtype']]
  elif isinstance(spec, tuple):
    v.tuple_value.SetInParent()
    v.tuple_value.values.extend(spec_value(s) for s in spec)
  else:
    v.list_value.SetInParent()
    v.list_value.values.extend(spec_value(s) for s in spec)
  return v


def jsonable(spec):
  if spec is None or isinstance(spec, dict):
    return spec
  return {'tuple' if isinstance(spec, tuple) else 'list': [jsonable(s) for s in spec]}


def main():
  rng = np.random.default_rng(3)
  tensors = [np.float32(1.5), np.arange(6, dtype=np.int32).reshape(2, 3), rng.integers(0, 256, (2, 4, 4, 1), dtype=np.uint8),
             np.array([2 ** 40, -7], np.int64), np.array([[True], [False]]), np.zeros((0, 3), np.float32),
             rng.normal(size=(3, 5)).astype(np.float32)]
  out = {'enum': {k: int(v) for k, v in DT.items()}, 'tensors': [], 'structures': []}
  for a in tensors:
    out['tensors'].append({'dtype': a.dtype.name, 'shape': list(a.shape), 'content': a.tobytes().hex(),
                           'hex': tensor_proto(a).SerializeToString().hex()})
  s1 = {'name': 'action', 'shape': [], 'dtype': 'int64'}
  s2 = {'name': None, 'shape': [84, 84, 4], 'dtype': 'uint8'}
  s3 = {'name': 'core', 'shape': [256], 'dtype': 'float32'}
  for spec in (s1, None, (s1, s2), [s3, (s1, (s2, s3))], (), [s2]):
    out['structures'].append({'spec': jsonable(spec), 'hex': spec_value(spec).SerializeToString().hex()})
  json.dump(out, open(os.path.join(HERE, 'tfproto_golden.json'), 'w'), indent=1)
  print('wrote tfproto_golden.json: %d tensors, %d structures' % (len(out['tensors']), len(out['structures'])))


if __name__ == '__main__':
  main()
