"""Wire-format mirror of TFQ's op inputs, built without protoc.

`tfq.proto.Program` and `tfq.proto.PauliSum` message classes are created at
import time from a hand-written FileDescriptorProto, so serialized bytes are
identical to what TFQ's serializer emits (reference:
tensorflow_quantum/core/proto/program.proto:20-161,
tensorflow_quantum/core/proto/pauli_sum.proto:20-35).

This module only *produces / inspects* the wire format for tests, generators
and the oracle; the product's decoder is the C++ one in csrc/wire.cc.
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=_F.LABEL_OPTIONAL, type_name=None,
           oneof_index=None):
    f = msg.field.add()
    f.name = name
    f.number = number
    f.type = ftype
    f.label = label
    if type_name:
        f.type_name = type_name
    if oneof_index is not None:
        f.oneof_index = oneof_index
    return f


def _build_pool():
    pool = descriptor_pool.DescriptorPool()

    # ---- program.proto -------------------------------------------------
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "tfqb/program.proto"
    fd.package = "tfq.proto"
    fd.syntax = "proto3"

    m = fd.message_type.add()
    m.name = "Program"
    m.oneof_decl.add().name = "program"
    _field(m, "language", 1, _F.TYPE_MESSAGE, type_name=".tfq.proto.Language")
    _field(m, "circuit", 2, _F.TYPE_MESSAGE, type_name=".tfq.proto.Circuit",
           oneof_index=0)
    _field(m, "schedule", 3, _F.TYPE_MESSAGE, type_name=".tfq.proto.Schedule",
           oneof_index=0)

    m = fd.message_type.add()
    m.name = "Circuit"
    e = m.enum_type.add()
    e.name = "SchedulingStrategy"
    v = e.value.add()
    v.name = "SCHEDULING_STRATEGY_UNSPECIFIED"
    v.number = 0
    v = e.value.add()
    v.name = "MOMENT_BY_MOMENT"
    v.number = 1
    _field(m, "scheduling_strategy", 1, _F.TYPE_ENUM,
           type_name=".tfq.proto.Circuit.SchedulingStrategy")
    _field(m, "moments", 2, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
           ".tfq.proto.Moment")

    m = fd.message_type.add()
    m.name = "Moment"
    _field(m, "operations", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
           ".tfq.proto.Operation")

    m = fd.message_type.add()
    m.name = "Schedule"
    _field(m, "scheduled_operations", 3, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
           ".tfq.proto.ScheduledOperation")

    m = fd.message_type.add()
    m.name = "ScheduledOperation"
    _field(m, "operation", 1, _F.TYPE_MESSAGE, type_name=".tfq.proto.Operation")
    _field(m, "start_time_picos", 2, _F.TYPE_INT64)

    m = fd.message_type.add()
    m.name = "Language"
    _field(m, "gate_set", 1, _F.TYPE_STRING)
    _field(m, "arg_function_language", 2, _F.TYPE_STRING)

    m = fd.message_type.add()
    m.name = "Operation"
    _field(m, "gate", 1, _F.TYPE_MESSAGE, type_name=".tfq.proto.Gate")
    _field(m, "args", 2, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
           ".tfq.proto.Operation.ArgsEntry")
    _field(m, "qubits", 3, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
           ".tfq.proto.Qubit")
    ent = m.nested_type.add()
    ent.name = "ArgsEntry"
    ent.options.map_entry = True
    _field(ent, "key", 1, _F.TYPE_STRING)
    _field(ent, "value", 2, _F.TYPE_MESSAGE, type_name=".tfq.proto.Arg")

    m = fd.message_type.add()
    m.name = "Gate"
    _field(m, "id", 1, _F.TYPE_STRING)

    m = fd.message_type.add()
    m.name = "Qubit"
    _field(m, "id", 2, _F.TYPE_STRING)

    m = fd.message_type.add()
    m.name = "Arg"
    m.oneof_decl.add().name = "arg"
    _field(m, "arg_value", 1, _F.TYPE_MESSAGE, type_name=".tfq.proto.ArgValue",
           oneof_index=0)
    _field(m, "symbol", 2, _F.TYPE_STRING, oneof_index=0)
    _field(m, "func", 3, _F.TYPE_MESSAGE, type_name=".tfq.proto.ArgFunction",
           oneof_index=0)

    m = fd.message_type.add()
    m.name = "ArgValue"
    m.oneof_decl.add().name = "arg_value"
    _field(m, "float_value", 1, _F.TYPE_FLOAT, oneof_index=0)
    _field(m, "bool_values", 2, _F.TYPE_MESSAGE,
           type_name=".tfq.proto.RepeatedBoolean", oneof_index=0)
    _field(m, "string_value", 3, _F.TYPE_STRING, oneof_index=0)
    _field(m, "double_value", 4, _F.TYPE_DOUBLE, oneof_index=0)

    m = fd.message_type.add()
    m.name = "RepeatedBoolean"
    _field(m, "values", 1, _F.TYPE_BOOL, _F.LABEL_REPEATED)

    m = fd.message_type.add()
    m.name = "ArgFunction"
    _field(m, "type", 1, _F.TYPE_STRING)
    _field(m, "args", 2, _F.TYPE_MESSAGE, _F.LABEL_REPEATED, ".tfq.proto.Arg")
    pool.Add(fd)

    # ---- pauli_sum.proto -----------------------------------------------
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "tfqb/pauli_sum.proto"
    fd.package = "tfq.proto"
    fd.syntax = "proto3"
    m = fd.message_type.add()
    m.name = "PauliSum"
    _field(m, "terms", 1, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
           ".tfq.proto.PauliTerm")
    m = fd.message_type.add()
    m.name = "PauliTerm"
    _field(m, "coefficient_real", 1, _F.TYPE_FLOAT)
    _field(m, "coefficient_imag", 2, _F.TYPE_FLOAT)
    _field(m, "paulis", 3, _F.TYPE_MESSAGE, _F.LABEL_REPEATED,
           ".tfq.proto.PauliQubitPair")
    m = fd.message_type.add()
    m.name = "PauliQubitPair"
    _field(m, "qubit_id", 1, _F.TYPE_STRING)
    _field(m, "pauli_type", 2, _F.TYPE_STRING)
    pool.Add(fd)
    return pool


_POOL = _build_pool()


def _cls(name):
    return message_factory.GetMessageClass(
        _POOL.FindMessageTypeByName("tfq.proto." + name))


Program = _cls("Program")
Circuit = _cls("Circuit")
Moment = _cls("Moment")
Operation = _cls("Operation")
Arg = _cls("Arg")
PauliSum = _cls("PauliSum")
PauliTerm = _cls("PauliTerm")
MOMENT_BY_MOMENT = 1
