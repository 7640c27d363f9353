"""Device-side evaluation of simple expressions: virtual columns, filters and selections of the chunk-feed loop.

Reference: every chunk, ``_BlockScope.evaluate`` runs Python ``eval(expression, expression_namespace, scope)`` over the numpy
blocks (packages/vaex-core/vaex/scopes.py:108-128); the pre-filter mask compresses every dependent column
(vaex/execution.py:516-522); selections are expressions evaluated the same way (:539, :551).  Here the expression is compiled
ONCE from its Python AST into a postfix program for ``b200_eval`` (csrc/expr.cu).  numpy decides the dtype of every node — the
compiler asks numpy itself (zero-length arrays, so NEP 50's weak Python scalars are honoured) — and the device performs exactly one
correctly rounded operation per node in that dtype, which makes the results bit-identical to the reference's numpy evaluation.

Supported: column names, numeric literals, ``+ - * /``, unary ``-``, ``abs() sqrt()``, ``< <= > >= == !=`` (not chained: numpy rejects `a < b < c` on arrays too), ``& | ~``
on booleans, ``.astype('dtype')`` and ``_ordinal_values(x, hash_map_unique)`` (vaex/functions.py:2454-2463).  Anything else raises
``NotImplementedError`` when the expression is compiled — never a silent CPU evaluation.
"""
import ast
import ctypes as C

import numpy as np

from . import _lib

(EX_INPUT, EX_CONST_F64, EX_CONST_I64, EX_ADD, EX_SUB, EX_MUL, EX_DIV, EX_NEG, EX_ABS, EX_SQRT, EX_LT, EX_LE, EX_GT, EX_GE, EX_EQ, EX_NE, EX_AND, EX_OR, EX_NOT,
 EX_CAST, EX_ORDINAL) = range(21)
EXC_F64, EXC_F32, EXC_I64, EXC_U64, EXC_BOOL = range(5)

_BINOPS = {ast.Add: (EX_ADD, np.add), ast.Sub: (EX_SUB, np.subtract), ast.Mult: (EX_MUL, np.multiply), ast.Div: (EX_DIV, np.true_divide)}
_CMPOPS = {ast.Lt: EX_LT, ast.LtE: EX_LE, ast.Gt: EX_GT, ast.GtE: EX_GE, ast.Eq: EX_EQ, ast.NotEq: EX_NE}


def _cls(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return EXC_F64
    if dtype == np.float32:
        return EXC_F32
    if dtype == np.bool_:
        return EXC_BOOL
    if dtype.kind == "i":
        return EXC_I64
    if dtype.kind == "u":
        return EXC_U64
    raise NotImplementedError(f"dtype {dtype} is not supported in device expressions")


class _Node:
    """value on the compile-time stack: a typed array expression, or a weak Python scalar (NEP 50) that takes the other
    operand's dtype"""

    def __init__(self, dtype=None, scalar=None):
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.scalar = scalar

    @property
    def weak(self):
        return self.dtype is None

    def probe(self):
        return self.scalar if self.weak else np.zeros(0, self.dtype)


class Program:
    """A compiled expression: postfix ops + the column names it reads + the hash maps it probes."""

    def __init__(self, expression, dtypes, variables=None):
        self.expression = str(expression)
        self.ops = []
        self.inputs = []  # column names, in input-index order
        self.sets = []    # ordered sets (superutils objects) probed by _ordinal_values
        self._dtypes = dtypes
        self._variables = dict(variables or {})
        tree = ast.parse(self.expression, mode="eval")
        top = self._emit(tree.body)
        if top.weak:  # a bare constant: give it numpy's default type
            top = self._materialise(top, np.result_type(top.scalar))
        self.dtype = top.dtype

    # ---- emit helpers -------------------------------------------------------------------------------------------------
    def _op(self, op, cls=0, arg=0, f=0.0, i=0):
        self.ops.append((op, cls, arg, float(f), int(i)))

    def _materialise(self, node, dtype):
        """push-time resolution of a weak scalar: it becomes a constant of `dtype`"""
        dtype = np.dtype(dtype)
        value = np.asarray(node.scalar).astype(dtype)[()]  # numpy's own conversion of the literal
        c = _cls(dtype)
        pos = node.slot
        if c in (EXC_F64, EXC_F32):
            self.ops[pos] = (EX_CONST_F64, c, 0, float(value), 0)
        else:
            self.ops[pos] = (EX_CONST_I64, c, 0, 0.0, int(value))
        return _Node(dtype)

    def _cast(self, node, dtype):
        dtype = np.dtype(dtype)
        if node.dtype != dtype:
            self._op(EX_CAST, _cls(node.dtype), _lib.DTYPE_CODE[dtype.name])
        return _Node(dtype)

    def _emit(self, n):
        if isinstance(n, ast.Name):
            if n.id not in self._dtypes:
                raise KeyError(f"column {n.id!r} not found (expression {self.expression!r})")
            dt = np.dtype(self._dtypes[n.id])
            if dt.kind in "mM" or dt.name not in _lib.DTYPE_CODE:
                raise NotImplementedError(f"column {n.id!r} has dtype {dt}: not supported in device expressions")
            if n.id not in self.inputs:
                self.inputs.append(n.id)
            self._op(EX_INPUT, _cls(dt), self.inputs.index(n.id))
            return _Node(dt)
        if isinstance(n, ast.Constant):
            if isinstance(n.value, bool) or not isinstance(n.value, (int, float)):
                raise NotImplementedError(f"literal {n.value!r} is not supported in device expressions")
            node = _Node(scalar=n.value)
            node.slot = len(self.ops)
            self._op(EX_CONST_F64)  # placeholder, typed when its partner is known
            return node
        if isinstance(n, ast.UnaryOp):
            if isinstance(n.op, ast.USub):
                if isinstance(n.operand, ast.Constant) and isinstance(n.operand.value, (int, float)) and not isinstance(n.operand.value, bool):
                    return self._emit(ast.copy_location(ast.Constant(-n.operand.value), n))
                a = self._emit(n.operand)
                if a.dtype.kind == "b":
                    raise NotImplementedError("numpy has no negative of a boolean")
                self._op(EX_NEG, _cls(a.dtype))
                return self._wrap(a.dtype)
            if isinstance(n.op, ast.Invert):
                a = self._emit(n.operand)
                if a.dtype != np.bool_:
                    raise NotImplementedError("~ is only supported on boolean expressions")
                self._op(EX_NOT, EXC_BOOL)
                return a
            if isinstance(n.op, ast.UAdd):
                return self._emit(n.operand)
            raise NotImplementedError(ast.dump(n.op))
        if isinstance(n, ast.BinOp):
            if isinstance(n.op, (ast.BitAnd, ast.BitOr)):
                a, b = self._emit(n.left), self._emit(n.right)
                if a.weak or b.weak or a.dtype != np.bool_ or b.dtype != np.bool_:
                    raise NotImplementedError("& and | are only supported on boolean expressions")
                self._op(EX_AND if isinstance(n.op, ast.BitAnd) else EX_OR, EXC_BOOL)
                return _Node(np.bool_)
            if type(n.op) not in _BINOPS:
                raise NotImplementedError(f"operator {type(n.op).__name__} is not supported in device expressions")
            opcode, ufunc = _BINOPS[type(n.op)]
            a = self._emit(n.left)
            mark = len(self.ops)
            b = self._emit(n.right)
            return self._binary(opcode, ufunc, a, b, mark)
        if isinstance(n, ast.Compare):
            result = None
            left = self._emit(n.left)
            if len(n.ops) != 1:
                raise NotImplementedError("chained comparisons are not supported in device expressions (numpy rejects them too)")
            mark = len(self.ops)
            right = self._emit(n.comparators[0])
            if type(n.ops[0]) not in _CMPOPS:
                raise NotImplementedError(type(n.ops[0]).__name__)
            a, b, common = self._promote(left, right, mark, np.less)
            self._op(_CMPOPS[type(n.ops[0])], _cls(common))
            result = _Node(np.bool_)
            return result
        if isinstance(n, ast.Call):
            return self._call(n)
        if isinstance(n, ast.Attribute):
            raise NotImplementedError("attribute access is only supported as .astype(...)")
        raise NotImplementedError(f"{type(n).__name__} is not supported in device expressions")

    def _wrap(self, dtype):
        """integer results narrower than 64 bits wrap like numpy's"""
        dtype = np.dtype(dtype)
        if dtype.kind in "iu" and dtype.itemsize < 8:
            self._op(EX_CAST, _cls(dtype), _lib.DTYPE_CODE[dtype.name])
        return _Node(dtype)

    def _promote(self, a, b, mark, ufunc):
        """numpy's input dtype for `ufunc(a, b)`; inserts the casts.  `mark` = index of the first op of b's code."""
        if a.weak and b.weak:
            raise NotImplementedError("constant folding of two literals is not supported: write the value")
        # ask numpy: the loop it selects tells both the computation dtype and the output dtype
        out = ufunc(a.probe(), b.probe())
        if ufunc in (np.less,):
            common = np.result_type(*(x.probe() for x in (a, b)))
        else:
            common = out.dtype
        if a.weak:
            a = self._materialise(a, common)
        elif a.dtype != common:  # cast a: its code ends right before `mark`
            self.ops.insert(mark, (EX_CAST, _cls(a.dtype), _lib.DTYPE_CODE[np.dtype(common).name], 0.0, 0))
            if getattr(b, "slot", None) is not None and b.weak:
                b.slot += 1
            a = _Node(common)
        if b.weak:
            b = self._materialise(b, common)
        elif b.dtype != common:
            b = self._cast(b, common)
        return a, b, np.dtype(common)

    def _binary(self, opcode, ufunc, a, b, mark):
        a, b, common = self._promote(a, b, mark, ufunc)
        if common == np.bool_:
            raise NotImplementedError("arithmetic on booleans is not supported in device expressions")
        if opcode == EX_DIV and common.kind not in "f":
            raise NotImplementedError("unexpected integer true-division loop")
        self._op(opcode, _cls(common))
        return self._wrap(common)

    def _call(self, n):
        # x.astype('float64')
        if isinstance(n.func, ast.Attribute) and n.func.attr == "astype":
            a = self._emit(n.func.value)
            if len(n.args) != 1 or not isinstance(n.args[0], ast.Constant) or not isinstance(n.args[0].value, str):
                raise NotImplementedError("astype takes one dtype string")
            dt = np.dtype(n.args[0].value)
            if dt.name not in _lib.DTYPE_CODE:
                raise NotImplementedError(f"astype({dt})")
            if a.weak:
                return self._materialise(a, dt)
            return self._cast(a, dt)
        if not isinstance(n.func, ast.Name):
            raise NotImplementedError("only plain function names can be called in device expressions")
        name = n.func.id
        if name in ("abs", "sqrt"):
            if len(n.args) != 1:
                raise TypeError(f"{name}() takes one argument")
            a = self._emit(n.args[0])
            if a.weak:
                raise NotImplementedError(f"{name}() of a literal: write the value")
            if name == "abs":
                if a.dtype.kind in "bu":
                    return a
                self._op(EX_ABS, _cls(a.dtype))
                return self._wrap(a.dtype)
            out = np.sqrt(a.probe()).dtype  # ints -> float64, float32 stays
            if out not in (np.float64, np.float32):
                raise NotImplementedError(f"sqrt of {a.dtype}")
            a = self._cast(a, out)
            self._op(EX_SQRT, _cls(out))
            return _Node(out)
        if name == "_ordinal_values":
            if len(n.args) != 2 or not isinstance(n.args[1], ast.Name):
                raise NotImplementedError("_ordinal_values(expression, hash_map_unique_variable)")
            hm = self._variables.get(n.args[1].id)
            if hm is None:
                raise KeyError(f"variable {n.args[1].id!r} not found")
            internal = getattr(hm, "_internal", hm)
            a = self._emit(n.args[0])
            key_dtype = np.dtype(internal._np_dtype())
            if a.weak:
                a = self._materialise(a, key_dtype)
            a = self._cast(a, key_dtype)  # the set is probed with keys of its own dtype
            if internal not in self.sets:
                self.sets.append(internal)
            self._op(EX_ORDINAL, _cls(key_dtype), self.sets.index(internal))
            # ordered_set::map_ordinal picks the narrowest signed type that holds the set (src/hash_primitives.hpp:611-623)
            count = len(internal)
            out = np.dtype("int8") if count < 2 ** 7 else np.dtype("int16") if count < 2 ** 15 else np.dtype("int32") if count < 2 ** 31 else np.dtype("int64")
            return self._cast(_Node(np.int64), out)
        raise NotImplementedError(f"function {name}() is not supported in device expressions")

    # ---- run ----------------------------------------------------------------------------------------------------------
    def c_ops(self):
        arr = (_lib.ExprOp * len(self.ops))()
        for k, (op, cls, arg, f, i) in enumerate(self.ops):
            arr[k].op, arr[k].cls, arr[k].arg, arr[k].f, arr[k].i = op, cls, arg, f, i
        return arr


class DeviceExpression:
    """A virtual column / filter / selection evaluated on the device chunk by chunk (the executor calls
    ``chunk(thread_index, i1, i2)`` on the worker's slot; the consumer runs on the same slot, i.e. the same stream)."""

    device_virtual = True

    def __init__(self, expression, columns, variables=None):
        self.expression = str(expression)
        self._columns = columns
        dtypes = {}
        for k, v in columns.items():
            if getattr(v, "device_virtual", False):
                dtypes[k] = v.dtype
            elif hasattr(v, "__cuda_array_interface__") and not isinstance(v, np.ndarray):
                dtypes[k] = np.dtype(v.__cuda_array_interface__["typestr"])
            else:
                dtypes[k] = v.dtype
        self.program = Program(expression, dtypes, variables)
        self.dtype = self.program.dtype
        self.columns = [columns[name] for name in self.program.inputs]  # the executor looks here to decide host / device feeding
        for name, col in zip(self.program.inputs, self.columns):
            if getattr(col, "device_virtual", False):
                raise NotImplementedError(f"{name!r} is itself a device-evaluated column: nest the expressions textually instead")
            if isinstance(col, np.ndarray) and np.ma.isMaskedArray(col):
                raise NotImplementedError(f"column {name!r} is masked: masked inputs are not supported in device expressions")
        self._ops = self.program.c_ops()
        self._buffers = {}
        self._ctx_cached = None

    def __len__(self):
        return len(self.columns[0]) if self.columns else 0

    @property
    def _ctx(self):
        if self._ctx_cached is None:
            self._ctx_cached = _lib.context()
        return self._ctx_cached

    def _buffer(self, thread_index, n):
        import torch
        buf = self._buffers.get(thread_index)
        nbytes = max(n, 1) * self.dtype.itemsize
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(nbytes + 16, dtype=torch.uint8, device=f"cuda:{self._ctx.device}")
            self._buffers[thread_index] = buf
        return buf

    def chunk(self, thread_index, i1, i2):
        n = int(i2 - i1)
        raw = self._buffer(thread_index, n)
        out = DeviceArray(raw.data_ptr(), n, self.dtype, keep=raw)
        if n == 0:
            return out
        nin = len(self.columns)
        inputs = (_lib.ExprInput * max(nin, 1))()
        keep, spaces = [], set()
        for k, col in enumerate(self.columns):
            part = col[i1:i2]
            if isinstance(part, np.ndarray) and not part.dtype.isnative:
                part = part.astype(part.dtype.newbyteorder("="))
            c = _lib.column(part)
            keep.append(c)
            spaces.add(c.memspace)
            inputs[k].data, inputs[k].dtype = c.ptr, c.code
        sets = (C.c_void_p * max(len(self.program.sets), 1))(*[s._h for s in self.program.sets])
        memspace = _lib.MEM_HOST if spaces == {_lib.MEM_HOST} else _lib.MEM_DEVICE if spaces <= {_lib.MEM_DEVICE} else _lib.MEM_MIXED
        _lib.check(_lib.lib().b200_eval(self._ctx._h, self._ctx.slot(thread_index), self._ops, len(self.program.ops), inputs, nin, sets, len(self.program.sets), n,
                                        memspace, _lib.DTYPE_CODE[self.dtype.name], raw.data_ptr()))
        return out

    def evaluate(self, i1=0, i2=None, thread_index=0):
        """host copy of the values (tests, small results)"""
        i2 = len(self) if i2 is None else i2
        return self.chunk(thread_index, i1, i2).to_numpy()


class DeviceArray:
    """A typed view of device memory (``__cuda_array_interface__``), what the device evaluators hand to the task parts."""

    def __init__(self, ptr, length, dtype, keep=None):
        self.ptr, self.length, self.dtype, self._keep = int(ptr), int(length), np.dtype(dtype), keep
        self.__cuda_array_interface__ = {"shape": (self.length,), "typestr": self.dtype.str, "data": (self.ptr, False), "version": 3, "strides": None}

    def __len__(self):
        return self.length

    def __getitem__(self, s):
        if not isinstance(s, slice):
            raise TypeError("only slices")
        a, b, step = s.indices(self.length)
        if step != 1:
            raise ValueError("only contiguous slices")
        return DeviceArray(self.ptr + a * self.dtype.itemsize, max(b - a, 0), self.dtype, keep=self._keep)

    def to_numpy(self):
        import torch
        _lib.context().sync()  # the producing kernel runs on a slot's stream, torch copies on its own
        out = np.empty(self.length, self.dtype)
        if self.length:
            t = torch.as_tensor(self, device="cuda")
            out[:] = t.cpu().numpy().view(self.dtype) if self.dtype.kind == "u" and self.dtype.itemsize > 1 else t.cpu().numpy()
        return out


def compact(thread_index, keep_mask, columns, ctx=None):
    """Filter compaction on the device (vaex/execution.py:516-522 `filter(v, filter_mask)`): returns (count, [device arrays]) with
    the rows of every column where keep_mask is non-zero, in order.  Columns may be host (numpy) or device arrays."""
    import torch
    ctx = ctx or _lib.context()
    n = len(keep_mask)
    cols = [_lib.column(c.astype(c.dtype.newbyteorder("=")) if isinstance(c, np.ndarray) and not c.dtype.isnative else c) for c in columns]
    km = _lib.mask_column(keep_mask)
    ncols = len(cols)
    bufs = [torch.empty(max(n, 1) * c.dtype.itemsize + 16, dtype=torch.uint8, device=f"cuda:{ctx.device}") for c in cols]
    cptr = (C.c_void_p * max(ncols, 1))(*[c.ptr for c in cols])
    optr = (C.c_void_p * max(ncols, 1))(*[b.data_ptr() for b in bufs])
    dts = (C.c_int32 * max(ncols, 1))(*[c.code for c in cols])
    count = C.c_int64(0)
    spaces = {c.memspace for c in cols} | {km.memspace}
    memspace = _lib.MEM_HOST if spaces == {_lib.MEM_HOST} else _lib.MEM_DEVICE if spaces == {_lib.MEM_DEVICE} else _lib.MEM_MIXED
    _lib.check(_lib.lib().b200_compact(ctx._h, ctx.slot(thread_index), km.ptr, ncols, cptr, dts, n, memspace, optr, C.byref(count)))
    return count.value, [DeviceArray(b.data_ptr(), count.value, c.dtype, keep=b) for b, c in zip(bufs, cols)]
