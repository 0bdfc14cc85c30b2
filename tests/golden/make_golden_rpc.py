"""Generates tests/golden/rpc_golden.json.  Run ONLY in the build container (where
/root/reference exists):   python tests/golden/make_golden_rpc.py

Takes the serialized FileDescriptorProto that protoc compiled from the reference's
grpc/service.proto (the `serialized_pb` literal of grpc/service_pb2.py, pulled out by AST: the
generated module itself does not import under protobuf 6), builds the message classes with the
protobuf runtime, and serialises sample messages of every type on the wire
(InitResponse / MethodOutputSignature / CallRequest / CallResponse).  The hand-written codec
of seed_rl_b200/grpc/ops.py is then checked against these bytes in both directions
(tests/test_rpc.py) -- the envelope of the reference's RPC surface is pinned to the
reference's own compiled schema, not to our reading of the .proto."""
import ast
import json
import os

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/grpc/service_pb2.py'


def serialized_descriptor():
  tree = ast.parse(open(REF).read())
  for node in ast.walk(tree):
    if isinstance(node, ast.keyword) and node.arg == 'serialized_pb':
      v = node.value
      if isinstance(v, ast.Call):          # _b('...')
        v = v.args[0]
      s = ast.literal_eval(v)
      return s.encode('latin1') if isinstance(s, str) else s
  raise KeyError('serialized_pb')


def main():
  fd = descriptor_pb2.FileDescriptorProto()
  fd.ParseFromString(serialized_descriptor())
  pool = descriptor_pool.DescriptorPool()
  pool.Add(fd)
  cls = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName('seed_rl.' + n))
  payloads = [b'', b'\x08\x01\x12\x00', bytes(range(256)) * 3, b'tensor-bytes-\xff\x00\x7f']
  out = {'descriptor_fields': {m.name: [(f.name, f.number, f.type, f.label) for f in m.field]
                               for m in fd.message_type}, 'cases': []}

  def add(kind, msg, **fields):
    out['cases'].append({'kind': kind, 'hex': msg.SerializeToString().hex(),
                         'fields': {k: (v.hex() if isinstance(v, bytes) else
                                        [x.hex() if isinstance(x, bytes) else x for x in v] if isinstance(v, list) else v)
                                    for k, v in fields.items()}})

  for fn, tensors in (('inference', payloads), ('', []), ('fünc', payloads[:1])):
    m = cls('CallRequest')(function=fn)
    m.tensor.extend(tensors)
    add('CallRequest', m, function=fn, tensor=list(tensors))
  for tensors, code, msg in ((payloads, 0, ''), ([], 3, 'Expects arg[0] to be int32 but float is provided'),
                             (payloads[2:], 13, 'Function bar not found'), ([], 0, '')):
    m = cls('CallResponse')(status_code=code, status_error_message=msg)
    m.tensor.extend(tensors)
    add('CallResponse', m, tensor=list(tensors), status_code=code, status_error_message=msg)
  m = cls('InitResponse')()
  sigs = [('inference', b'\x9a\x02\x05spec1'), ('other_fn', b''), ('', b'\x01\x02')]
  for n, spec in sigs:
    s = m.method_output_signature.add()
    s.name = n
    s.output_specs = spec
  add('InitResponse', m, names=[n for n, _ in sigs], specs=[s for _, s in sigs])
  add('InitRequest', cls('InitRequest')())
  json.dump(out, open(os.path.join(HERE, 'rpc_golden.json'), 'w'), indent=1)
  print('wrote rpc_golden.json: %d cases' % len(out['cases']))


if __name__ == '__main__':
  main()
