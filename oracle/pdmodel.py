"""ORACLE (test infrastructure, never shipped): schema-less reader for Paddle
`.pdmodel` (ProgramDesc protobuf) and `.pdiparams` (raw LoDTensor stream).

Field numbers follow the reference's vendored schema
`include/paddle_inference/internal/framework.pb.h`:
  ProgramDesc{blocks=1}; BlockDesc{idx=1,parent_idx=2,vars=3,ops=4};
  OpDesc{inputs=1,outputs=2,type=3,attrs=4}; OpDesc.Var{parameter=1,arguments=2};
  OpDesc.Attr{name=1,type=2,i=3,f=4,s=5,ints=6,floats=7,strings=8,b=10,bools=11,
              block_idx=12,l=13,longs=15,float64s=16,float64=19};
  VarDesc{name=1,type=2,persistable=3}; VarType{type=1,lod_tensor=3};
  LoDTensorDesc{tensor=1}; TensorDesc{data_type=1,dims=2}.
The reference itself hands these files to `config.SetModel` (src/ocr_det.cpp:46).
"""
from __future__ import annotations
import struct
from dataclasses import dataclass, field
import numpy as np


def _varint(buf, pos):
    r = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        r |= (b & 0x7F) << shift
        if not (b & 0x80):
            return r, pos
        shift += 7


def _fields(buf):
    """Yield (field_no, wire_type, value) for one message."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]; pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _s64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _s32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v >= (1 << 31) else v


def _packed_varints(wt, v):
    if wt == 0:
        return [_s64(v)]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_s64(x))
    return out


@dataclass
class Op:
    type: str
    inputs: dict = field(default_factory=dict)
    outputs: dict = field(default_factory=dict)
    attrs: dict = field(default_factory=dict)

    def i(self, k):  # single input name
        return self.inputs[k][0]

    def o(self, k):
        return self.outputs[k][0]


@dataclass
class Var:
    name: str
    persistable: bool = False
    dtype: int = -1
    dims: tuple = ()
    vtype: int = -1


@dataclass
class Program:
    vars: dict
    ops: list


def _parse_attr(buf):
    name, val = None, None
    ints, floats, strings, bools, longs, f64s = [], [], [], [], [], []
    atype = None
    for fno, wt, v in _fields(buf):
        if fno == 1: name = bytes(v).decode()
        elif fno == 2: atype = v
        elif fno == 3: val = ('i', v)
        elif fno == 4: val = ('f', struct.unpack('<f', bytes(v))[0])
        elif fno == 5: val = ('s', bytes(v).decode('utf-8', 'replace'))
        elif fno == 6: ints += _packed_varints(wt, v)
        elif fno == 7:
            floats += list(struct.unpack(f'<{len(v)//4}f', bytes(v))) if wt == 2 else [struct.unpack('<f', bytes(v))[0]]
        elif fno == 8: strings.append(bytes(v).decode('utf-8', 'replace'))
        elif fno == 10: val = ('b', bool(v))
        elif fno == 11: bools += [bool(x) for x in _packed_varints(wt, v)]
        elif fno == 12: val = ('block', v)
        elif fno == 13: val = ('l', _s64(v))
        elif fno == 15: longs += _packed_varints(wt, v)
        elif fno == 19: val = ('d', struct.unpack('<d', bytes(v))[0])
    # AttrType enum: INT=0 FLOAT=1 STRING=2 INTS=3 FLOATS=4 STRINGS=5 BOOLEAN=6
    # BOOLEANS=7 BLOCK=8 LONG=9 BLOCKS=10 LONGS=11 FLOAT64S=12 VAR=13 VARS=14 FLOAT64=15
    if atype == 0:
        out = _s32(val[1]) if val else 0
    elif atype == 1: out = val[1] if val else 0.0
    elif atype == 2: out = val[1] if val else ''
    elif atype == 3: out = [_s32(x) for x in ints]
    elif atype == 4: out = floats
    elif atype == 5: out = strings
    elif atype == 6: out = val[1] if val else False
    elif atype == 7: out = bools
    elif atype == 9: out = val[1] if val else 0
    elif atype == 11: out = longs
    elif atype == 15: out = val[1] if val else 0.0
    else: out = val[1] if val else None
    return name, out


def _parse_opvar(buf):
    param, args = None, []
    for fno, wt, v in _fields(buf):
        if fno == 1: param = bytes(v).decode()
        elif fno == 2: args.append(bytes(v).decode())
    return param, args


def _parse_op(buf):
    op = Op(type='')
    for fno, wt, v in _fields(buf):
        if fno == 3: op.type = bytes(v).decode()
        elif fno == 1:
            p, a = _parse_opvar(v); op.inputs[p] = a
        elif fno == 2:
            p, a = _parse_opvar(v); op.outputs[p] = a
        elif fno == 4:
            n, val = _parse_attr(v)
            if n not in ('op_callstack', 'op_namescope', 'op_role', 'op_role_var', 'op_device', 'with_quant_attr'):
                op.attrs[n] = val
    return op


def _parse_tensor_desc(buf):
    dtype, dims = -1, []
    for fno, wt, v in _fields(buf):
        if fno == 1: dtype = v
        elif fno == 2: dims += _packed_varints(wt, v)
    return dtype, tuple(dims)


def _parse_var(buf):
    var = Var(name='')
    for fno, wt, v in _fields(buf):
        if fno == 1: var.name = bytes(v).decode()
        elif fno == 3: var.persistable = bool(v)
        elif fno == 2:  # VarType
            for f2, w2, v2 in _fields(v):
                if f2 == 1: var.vtype = v2
                elif f2 == 3:  # LoDTensorDesc
                    for f3, w3, v3 in _fields(v2):
                        if f3 == 1: var.dtype, var.dims = _parse_tensor_desc(v3)
    return var


def load_program(path) -> Program:
    buf = memoryview(open(path, 'rb').read())
    vars_, ops = {}, []
    for fno, wt, v in _fields(buf):
        if fno == 1:  # BlockDesc (single block in all three shipped graphs)
            for f2, w2, v2 in _fields(v):
                if f2 == 3:
                    var = _parse_var(v2); vars_[var.name] = var
                elif f2 == 4:
                    ops.append(_parse_op(v2))
    return Program(vars_, ops)


def param_names(prog: Program):
    """Persistable vars in the order `.pdiparams` stores them (ascending name)."""
    return sorted(n for n, v in prog.vars.items()
                  if v.persistable and n not in ('feed', 'fetch') and v.vtype == 7)


def load_params(prog: Program, path) -> dict:
    """.pdiparams record: u32 0 | u64 lod_level | u32 0 | i32 desc_len | TensorDesc | fp32 data."""
    buf = memoryview(open(path, 'rb').read())
    pos, out = 0, {}
    for name in param_names(prog):
        _ver, = struct.unpack_from('<I', buf, pos); pos += 4
        lod, = struct.unpack_from('<Q', buf, pos); pos += 8
        for _ in range(lod):
            sz, = struct.unpack_from('<Q', buf, pos); pos += 8 + sz
        _tver, = struct.unpack_from('<I', buf, pos); pos += 4
        dl, = struct.unpack_from('<i', buf, pos); pos += 4
        dtype, dims = _parse_tensor_desc(buf[pos:pos + dl]); pos += dl
        assert dtype == 5, (name, dtype)
        assert tuple(dims) == tuple(prog.vars[name].dims), (name, dims, prog.vars[name].dims)
        n = int(np.prod(dims)) if dims else 1
        out[name] = np.frombuffer(buf, dtype='<f4', count=n, offset=pos).reshape(dims).copy()
        pos += 4 * n
    assert pos == len(buf), (pos, len(buf))
    return out


def _enc_varint(x):
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b); return bytes(out)


def save_params(prog: Program, params: dict, path):
    with open(path, 'wb') as f:
        for name in param_names(prog):
            a = np.ascontiguousarray(params[name], dtype='<f4')
            desc = b'\x08\x05' + b''.join(b'\x10' + _enc_varint(int(d)) for d in a.shape)
            f.write(struct.pack('<IQIi', 0, 0, 0, len(desc)))
            f.write(desc)
            f.write(a.tobytes())
