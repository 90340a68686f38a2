"""Minimal CPython-3.8 .pyc reader + symbolic decompiler for straight-line functions.

The reference ships two modules only as stale bytecode (__pycache__/physics_functions.cpython-38.pyc and
l_bfgs_b_optimizer.cpython-38.pyc).  Python 3.12's marshal cannot load 3.8 code objects, so this
reads the marshal stream by hand and replays each function's bytecode on a symbolic stack, printing one source-like
statement per STORE / RETURN.  Used once, in the build container, to recover the formulas of `Boussinesq` that
oracle/boussinesq_oracle.py re-types.      python tools/pyc38_decompile.py /root/reference/__pycache__/physics_functions.cpython-38.pyc
"""
import struct
import sys

OP = {1: "POP_TOP", 2: "ROT_TWO", 3: "ROT_THREE", 4: "DUP_TOP", 5: "DUP_TOP_TWO", 9: "NOP", 10: "UNARY_POSITIVE",
      11: "UNARY_NEGATIVE", 12: "UNARY_NOT", 19: "BINARY_POWER", 20: "BINARY_MULTIPLY", 22: "BINARY_MODULO", 23: "BINARY_ADD",
      24: "BINARY_SUBTRACT", 25: "BINARY_SUBSCR", 26: "BINARY_FLOOR_DIVIDE", 27: "BINARY_TRUE_DIVIDE", 55: "INPLACE_ADD",
      56: "INPLACE_SUBTRACT", 57: "INPLACE_MULTIPLY", 29: "INPLACE_TRUE_DIVIDE", 67: "INPLACE_POWER", 60: "STORE_SUBSCR",
      68: "GET_ITER", 83: "RETURN_VALUE", 87: "POP_BLOCK", 90: "STORE_NAME", 92: "UNPACK_SEQUENCE", 93: "FOR_ITER",
      95: "STORE_ATTR", 97: "STORE_GLOBAL", 100: "LOAD_CONST", 101: "LOAD_NAME", 102: "BUILD_TUPLE", 103: "BUILD_LIST",
      105: "BUILD_MAP", 106: "LOAD_ATTR", 107: "COMPARE_OP", 108: "IMPORT_NAME", 109: "IMPORT_FROM", 110: "JUMP_FORWARD",
      113: "JUMP_ABSOLUTE", 114: "POP_JUMP_IF_FALSE", 115: "POP_JUMP_IF_TRUE", 116: "LOAD_GLOBAL", 124: "LOAD_FAST",
      125: "STORE_FAST", 126: "DELETE_FAST", 131: "CALL_FUNCTION", 132: "MAKE_FUNCTION", 133: "BUILD_SLICE",
      141: "CALL_FUNCTION_KW", 144: "EXTENDED_ARG", 156: "BUILD_CONST_KEY_MAP", 160: "LOAD_METHOD", 161: "CALL_METHOD"}
BIN = {"BINARY_POWER": "**", "BINARY_MULTIPLY": "*", "BINARY_MODULO": "%", "BINARY_ADD": "+", "BINARY_SUBTRACT": "-",
       "BINARY_FLOOR_DIVIDE": "//", "BINARY_TRUE_DIVIDE": "/", "INPLACE_ADD": "+", "INPLACE_SUBTRACT": "-",
       "INPLACE_MULTIPLY": "*", "INPLACE_TRUE_DIVIDE": "/", "INPLACE_POWER": "**"}


class Code:
    pass


class Reader:
    def __init__(self, data):
        self.d, self.p, self.refs = data, 0, []

    def u8(self):
        v = self.d[self.p]
        self.p += 1
        return v

    def i32(self):
        v = struct.unpack_from("<i", self.d, self.p)[0]
        self.p += 4
        return v

    def raw(self, n):
        v = self.d[self.p:self.p + n]
        self.p += n
        return v

    def obj(self):
        t = self.u8()
        flag, t = t & 0x80, chr(t & 0x7F)
        idx = None
        if flag:
            idx = len(self.refs)
            self.refs.append(None)
        v = self._obj(t)
        if flag:
            self.refs[idx] = v
        return v

    def _obj(self, t):
        if t == "0":
            return None
        if t == "N":
            return None
        if t == "F":
            return False
        if t == "T":
            return True
        if t == ".":
            return Ellipsis
        if t == "i":
            return self.i32()
        if t == "l":
            n = self.i32()
            digits = [struct.unpack_from("<H", self.raw(2))[0] for _ in range(abs(n))]
            v = sum(dg << (15 * k) for k, dg in enumerate(digits))
            return -v if n < 0 else v
        if t == "g":
            return struct.unpack("<d", self.raw(8))[0]
        if t == "f":
            return float(self.raw(self.u8()).decode())
        if t in "st":
            return self.raw(self.i32())
        if t in "uaA":
            return self.raw(self.i32()).decode("utf8", "replace")
        if t in "zZ":
            return self.raw(self.u8()).decode("latin1")
        if t == "(":
            return tuple(self.obj() for _ in range(self.i32()))
        if t == ")":
            return tuple(self.obj() for _ in range(self.u8()))
        if t == "[":
            return [self.obj() for _ in range(self.i32())]
        if t == "r":
            return self.refs[self.i32()]
        if t == "c":
            c = Code()
            (c.argcount, c.posonly, c.kwonly, c.nlocals, c.stacksize, c.flags) = (self.i32() for _ in range(6))
            c.code = self.obj()
            c.consts = self.obj()
            c.names = self.obj()
            c.varnames = self.obj()
            c.freevars = self.obj()
            c.cellvars = self.obj()
            c.filename = self.obj()
            c.name = self.obj()
            c.firstlineno = self.i32()
            c.lnotab = self.obj()
            return c
        raise ValueError(f"marshal type {t!r} at {self.p}")


class E(str):
    """expression text; .atom = needs no parentheses when used as an operand"""
    atom = True


def X(text, atom=True):
    e = E(text)
    e.atom = atom
    return e


def prec(s):
    return s if getattr(s, "atom", False) else f"({s})"


def decompile(c, out):
    code = c.code
    stack = []
    i, ext = 0, 0
    out.append(f"def {c.name}({', '.join(c.varnames[:c.argcount])}):   # {c.filename}:{c.firstlineno}")
    while i < len(code):
        op, arg = code[i], code[i + 1] | ext
        i += 2
        name = OP.get(op, f"OP{op}")
        ext = 0
        if name == "EXTENDED_ARG":
            ext = arg << 8
            continue
        if name == "LOAD_CONST":
            v = c.consts[arg]
            stack.append(X(f"<code {v.name}>" if isinstance(v, Code) else repr(v), not (isinstance(v, (int, float)) and not isinstance(v, bool) and v < 0)))
        elif name in ("LOAD_FAST",):
            stack.append(X(c.varnames[arg]))
        elif name in ("LOAD_GLOBAL", "LOAD_NAME"):
            stack.append(X(c.names[arg]))
        elif name in ("LOAD_ATTR", "LOAD_METHOD"):
            stack.append(X(f"{prec(stack.pop())}.{c.names[arg]}"))
        elif name in BIN:
            b, a = stack.pop(), stack.pop()
            stack.append(X(f"{prec(a)} {BIN[name]} {prec(b)}", False))
        elif name == "UNARY_NEGATIVE":
            stack.append(X(f"-{prec(stack.pop())}", False))
        elif name == "BINARY_SUBSCR":
            b, a = stack.pop(), stack.pop()
            stack.append(X(f"{prec(a)}[{b}]"))
        elif name == "BUILD_SLICE":
            parts = [stack.pop() for _ in range(arg)][::-1]
            stack.append(X(":".join("" if p_ == "None" else p_ for p_ in parts)))
        elif name in ("BUILD_TUPLE", "BUILD_LIST"):
            parts = [stack.pop() for _ in range(arg)][::-1]
            stack.append(X(", ".join(parts), False) if name == "BUILD_TUPLE" else X("[" + ", ".join(parts) + "]"))
        elif name in ("CALL_FUNCTION", "CALL_METHOD"):
            args = [stack.pop() for _ in range(arg)][::-1]
            f = stack.pop()
            stack.append(X(f"{f}({', '.join(args)})"))
        elif name == "CALL_FUNCTION_KW":
            kw = eval(stack.pop())
            args = [stack.pop() for _ in range(arg)][::-1]
            f = stack.pop()
            npos = len(args) - len(kw)
            stack.append(X(f"{f}({', '.join(args[:npos] + [f'{k}={v}' for k, v in zip(kw, args[npos:])])})"))
        elif name in ("STORE_FAST", "STORE_NAME", "STORE_GLOBAL"):
            tgt = c.varnames[arg] if name == "STORE_FAST" else c.names[arg]
            out.append(f"    {tgt} = {stack.pop()}")
        elif name == "UNPACK_SEQUENCE":
            v = stack.pop()
            for k in range(arg - 1, -1, -1):
                stack.append(X(f"{prec(v)}[{k}]"))
        elif name == "RETURN_VALUE":
            out.append(f"    return {stack.pop()}")
        elif name == "POP_TOP":
            out.append(f"    {stack.pop()}")
        elif name == "DUP_TOP":
            stack.append(stack[-1])
        elif name == "ROT_TWO":
            stack[-1], stack[-2] = stack[-2], stack[-1]
        elif name == "MAKE_FUNCTION":
            stack.pop()
            fn = stack.pop()
            stack.append(fn)
        elif name in ("IMPORT_NAME",):
            stack.pop(); stack.pop()
            stack.append(f"__import__({c.names[arg]!r})")
        elif name == "IMPORT_FROM":
            stack.append(f"{stack[-1]}.{c.names[arg]}")
        elif name == "COMPARE_OP":
            b, a = stack.pop(), stack.pop()
            stack.append(f"{a} <cmp{arg}> {b}")
        else:
            out.append(f"    # {name} {arg}   (stack: {stack[-3:]})")
    for v in c.consts:
        if isinstance(v, Code):
            out.append("")
            decompile(v, out)


def main():
    data = open(sys.argv[1], "rb").read()
    r = Reader(data[16:])
    top = r.obj()
    out = []
    decompile(top, out)
    print("\n".join(out))


if __name__ == "__main__":
    main()
