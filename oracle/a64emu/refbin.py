"""refbin.py — loads the reference's shipped Mach-O arm64 binary into the a64emu interpreter and calls its functions.

TEST INFRASTRUCTURE (oracle/).  `/root/reference/test-dist/xfg-stark-cli` is the only executable form of the reference's hot path
(`air.prove(trace)`, src/burn_mint_prover.rs:124 -> winter_prover::Prover::prove of winterfell 0.8.3) and it cannot run on this x86-64
Linux box.  This module parses the Mach-O (segments, symbol table, indirect symbols -> `__stubs` imports, thread-local descriptors), maps it
at its preferred address (slide 0, so no rebasing is needed), binds every libSystem import either to a native of the interpreter
(malloc/free/memcpy/...) or to a Python callback, and runs Rust functions by symbol name with AAPCS64 arguments.

The binary is opened read-only and never modified; nothing is copied from it into the repo.  Golden vectors produced through this module
are committed under tests/golden/ together with the script that made them (oracle/a64emu/make_reference_vectors.py).
"""
import ctypes as C
import os
import re
import struct
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
DEFAULT_BINARY = "/root/reference/test-dist/xfg-stark-cli"
LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "liba64emu.so")

R_DONE, R_HOOK, R_UNKNOWN, R_FAULT, R_TRAP, R_LIMIT = range(6)
NATIVES = {"_malloc": 2, "_free": 3, "_calloc": 4, "_realloc": 5, "_posix_memalign": 6, "_memcpy": 7, "_memmove": 8, "_memset": 9, "_memcmp": 10,
           "_bzero": 11, "_strlen": 12}
RET0 = 13
TLV = 14
RET0_IMPORTS = {"_pthread_mutex_init", "_pthread_mutex_lock", "_pthread_mutex_unlock", "_pthread_mutex_trylock", "_pthread_mutex_destroy",
                "_pthread_mutexattr_init", "_pthread_mutexattr_settype", "_pthread_mutexattr_destroy", "__tlv_atexit", "_sigaction", "_sigaltstack",
                "_signal", "_pthread_setname_np", "_munmap", "_mprotect", "_close", "_isatty", "_dispatch_release"}

STACK_BASE, STACK_SIZE = 0x7FF0000000, 64 << 20
HEAP_BASE, HEAP_SIZE = 0x200000000, 24 << 30
STOP_PC = 0xDEAD0000


def build_lib():
    src = [os.path.join(HERE, "a64emu.cpp"), os.path.join(HERE, "a64simd.inc")]
    if os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB_PATH, src[0]])
    return LIB_PATH


def demangle(s):
    """legacy Rust mangling (_ZN...E) -> path; good enough to look symbols up by name"""
    if not s.startswith("__ZN"):
        return s
    t = s[4:]; parts = []
    while t and t[0].isdigit():
        m = re.match(r"(\d+)", t); n = int(m.group(1)); t = t[len(m.group(1)):]
        parts.append(t[:n]); t = t[n:]
    out = "::".join(parts)
    for a, b in (("$LT$", "<"), ("$GT$", ">"), ("$u20$", " "), ("$C$", ","), ("$RF$", "&"), ("$LP$", "("), ("$RP$", ")"), ("$u7b$", "{"), ("$u7d$", "}"),
                 ("$BP$", "*"), ("$u5b$", "["), ("$u5d$", "]"), ("$u3b$", ";"), ("..", "::")):
        out = out.replace(a, b)
    return out[1:] if out.startswith("_<") else out


class MachO:
    """the parts of a 64-bit little-endian Mach-O this project needs"""

    def __init__(self, path=DEFAULT_BINARY):
        self.path = path
        with open(path, "rb") as fh:          # read-only
            f = self.data = fh.read()
        magic, cputype, _, filetype, ncmds = struct.unpack_from("<IiiII", f, 0)
        if magic != 0xFEEDFACF or cputype != 0x0100000C:
            raise ValueError("not a Mach-O arm64 image")
        self.segments, self.sections = [], {}
        off = 32
        for _ in range(ncmds):
            cmd, sz = struct.unpack_from("<II", f, off)
            if cmd == 0x19:
                name = f[off + 8:off + 24].rstrip(b"\0").decode()
                vmaddr, vmsize, fileoff, filesize, _, _, nsects, _ = struct.unpack_from("<QQQQiiII", f, off + 24)
                self.segments.append((name, vmaddr, vmsize, fileoff, filesize))
                so = off + 72
                for _s in range(nsects):
                    sn = f[so:so + 16].rstrip(b"\0").decode(); sg = f[so + 16:so + 32].rstrip(b"\0").decode()
                    addr, size, offset, _, _, _, flags, r1, r2, _ = struct.unpack_from("<QQIIIIIIII", f, so + 32)
                    self.sections[(sg, sn)] = dict(addr=addr, size=size, offset=offset, flags=flags, r1=r1, r2=r2)
                    so += 80
            elif cmd == 0x2:
                self.symoff, self.nsyms, self.stroff, self.strsize = struct.unpack_from("<IIII", f, off + 8)
            elif cmd == 0xB:
                d = struct.unpack_from("<18I", f, off + 8); self.indoff, self.nind = d[12], d[13]
            elif cmd == 0x80000022:
                self.dyld_info = struct.unpack_from("<10I", f, off + 8)
            off += sz
        self.symbols = []                      # (mangled, demangled, type, sect, value)
        for i in range(self.nsyms):
            strx, typ, sect, desc, val = struct.unpack_from("<IBBHQ", f, self.symoff + 16 * i)
            e = f.index(b"\0", self.stroff + strx); nm = f[self.stroff + strx:e].decode()
            self.symbols.append((nm, demangle(nm), typ, sect, val))
        self.indirect = struct.unpack_from("<%dI" % self.nind, f, self.indoff)
        self.funcs = sorted((v, d, m) for m, d, t, s, v in self.symbols if (t & 0xE) == 0xE and s == 1)
        self._addrs = [a for a, _, _ in self.funcs]
        st = self.sections[("__TEXT", "__stubs")]
        self.stubs = {st["addr"] + 12 * k: self.symbols[self.indirect[st["r1"] + k]][0] for k in range(st["size"] // 12)}

    def find(self, pattern, index=None):
        """address(es) of functions whose demangled name matches the regular expression; index picks one monomorphisation (address order)"""
        rx = re.compile(pattern)
        hits = [(a, d) for a, d, _ in self.funcs if rx.search(d)]
        seen, uniq = set(), []
        for a, d in hits:
            if a not in seen:
                seen.add(a); uniq.append((a, d))
        if index is None:
            return uniq
        return uniq[index][0]

    def func_size(self, addr):
        import bisect
        i = bisect.bisect_right(self._addrs, addr)
        text = self.sections[("__TEXT", "__text")]
        end = self._addrs[i] if i < len(self._addrs) else text["addr"] + text["size"]
        return end - addr

    def symbol_at(self, pc):
        import bisect
        i = bisect.bisect_right(self._addrs, pc) - 1
        if i < 0:
            return "?"
        a, d, _ = self.funcs[i]
        return f"{d}+0x{pc - a:x}"

    def read(self, vmaddr, n):
        for name, va, vs, fo, fs in self.segments:
            if va <= vmaddr < va + fs:
                return self.data[fo + vmaddr - va:fo + vmaddr - va + n]
        raise KeyError(hex(vmaddr))

    def binds(self):
        """non-lazy bind records (segment index, offset, symbol) from LC_DYLD_INFO"""
        _, _, boff, bsize = self.dyld_info[0], self.dyld_info[1], self.dyld_info[2], self.dyld_info[3]
        p = boff; end = boff + bsize; f = self.data
        out = []; seg = 0; offs = 0; sym = ""; typ = 1

        def uleb():
            nonlocal p
            r = 0; sh = 0
            while True:
                b = f[p]; p += 1; r |= (b & 0x7F) << sh; sh += 7
                if not b & 0x80:
                    return r
        while p < end:
            b = f[p]; p += 1; op = b & 0xF0; imm = b & 0x0F
            if op == 0x00:
                break
            elif op == 0x10 or op == 0x30:
                pass
            elif op == 0x20:
                uleb()
            elif op == 0x40:
                e = f.index(b"\0", p); sym = f[p:e].decode(); p = e + 1
            elif op == 0x50:
                typ = imm
            elif op == 0x60:
                uleb()
            elif op == 0x70:
                seg = imm; offs = uleb()
            elif op == 0x80:
                offs += uleb()
            elif op == 0x90:
                out.append((seg, offs, sym)); offs += 8
            elif op == 0xA0:
                out.append((seg, offs, sym)); offs += 8 + uleb()
            elif op == 0xB0:
                out.append((seg, offs, sym)); offs += 8 + imm * 8
            elif op == 0xC0:
                cnt = uleb(); skip = uleb()
                for _ in range(cnt):
                    out.append((seg, offs, sym)); offs += 8 + skip
        return out


class EmuError(RuntimeError):
    pass


class RefBinary:
    """the reference binary loaded in the interpreter"""

    def __init__(self, path=DEFAULT_BINARY, verbose=False):
        self.m = MachO(path)
        L = self.L = C.CDLL(build_lib())
        L.emu_create.restype = C.c_void_p
        for name, args, res in (("emu_map", [C.c_void_p, C.c_uint64, C.c_uint64], C.c_int), ("emu_set_heap", [C.c_void_p, C.c_uint64, C.c_uint64], C.c_int),
                                ("emu_write", [C.c_void_p, C.c_uint64, C.c_char_p, C.c_uint64], C.c_int), ("emu_read", [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64], C.c_int),
                                ("emu_malloc", [C.c_void_p, C.c_uint64], C.c_uint64), ("emu_set_text", [C.c_void_p, C.c_uint64, C.c_uint64], None),
                                ("emu_hook", [C.c_void_p, C.c_uint64, C.c_int], C.c_int), ("emu_regs", [C.c_void_p], C.POINTER(C.c_uint64)),
                                ("emu_get", [C.c_void_p, C.c_int], C.c_uint64), ("emu_set", [C.c_void_p, C.c_int, C.c_uint64], None),
                                ("emu_vreg", [C.c_void_p, C.c_int], C.POINTER(C.c_uint64)), ("emu_run", [C.c_void_p, C.c_uint64], C.c_int),
                                ("emu_return_from_hook", [C.c_void_p], None), ("emu_step_over", [C.c_void_p], C.c_int), ("emu_destroy", [C.c_void_p], None)):
            fn = getattr(L, name); fn.argtypes = args; fn.restype = res
        self.e = L.emu_create()
        self.verbose = verbose
        self.output = []            # bytes the guest wrote to fd 1 / 2
        self.py_hooks = {}          # address -> callable(self) -> None (sets x0) ; return "continue" to execute the hooked instruction
        m = self.m
        lo = min(va for n, va, vs, fo, fs in m.segments if n not in ("__PAGEZERO", "__LINKEDIT"))
        hi = max(va + vs for n, va, vs, fo, fs in m.segments if n not in ("__PAGEZERO", "__LINKEDIT"))
        self.image = (lo, hi - lo)
        assert L.emu_map(self.e, lo, hi - lo) == 0
        for n, va, vs, fo, fs in m.segments:
            if n in ("__PAGEZERO", "__LINKEDIT") or not fs:
                continue
            self.write(va, m.data[fo:fo + fs])
        tx = [s for s in m.segments if s[0] == "__TEXT"][0]
        L.emu_set_text(self.e, tx[1], tx[2])
        assert L.emu_map(self.e, STACK_BASE, STACK_SIZE) == 0
        assert L.emu_set_heap(self.e, HEAP_BASE, HEAP_SIZE) == 0
        L.emu_set(self.e, 9, STOP_PC)
        self.x = L.emu_regs(self.e)
        # imports
        for addr, name in m.stubs.items():
            if name in NATIVES:
                L.emu_hook(self.e, addr, NATIVES[name])
            elif name in RET0_IMPORTS:
                L.emu_hook(self.e, addr, RET0)
            else:
                bound = getattr(self, "imp" + name, None)
                L.emu_hook(self.e, addr, 1); self.py_hooks[addr] = (lambda rb, f=bound: f()) if bound else self._unimplemented(name)
        # data binds (non-lazy): ___stack_chk_guard and the thread-local bootstrap thunk
        segs = m.segments
        self.guard = self.malloc(16); self.write(self.guard, struct.pack("<Q", 0x595E9FBD94FDA766))
        helper = m.sections[("__TEXT", "__stub_helper")]["addr"]      # never executed (stubs are hooked): reused as the TLV thunk address
        self.tlv_thunk = helper
        L.emu_hook(self.e, helper, TLV)
        for seg, offs, sym in m.binds():
            target = segs[seg][1] + offs
            if sym == "___stack_chk_guard":
                self.write(target, struct.pack("<Q", self.guard))
            elif sym == "__tlv_bootstrap":
                self.write(target, struct.pack("<Q", helper))
            elif sym == "dyld_stub_binder":
                pass
            else:
                raise EmuError(f"unhandled data bind {sym}")
        # thread-local storage: template = __thread_data followed by __thread_bss
        td = m.sections.get(("__DATA", "__thread_data")); tb = m.sections.get(("__DATA", "__thread_bss"))
        if td:
            total = (tb["addr"] + tb["size"] - td["addr"]) if tb else td["size"]
            self.tls = self.malloc(total + 64); self.write(self.tls, m.read(td["addr"], td["size"]) + b"\0" * (total - td["size"]))
            L.emu_set(self.e, 10, self.tls)
        self.errno_ptr = self.malloc(16); self.write(self.errno_ptr, b"\0" * 16)

    # ---- memory helpers ----
    def write(self, addr, data):
        if self.L.emu_write(self.e, addr, bytes(data), len(data)) != 0:
            raise EmuError(f"write fault at {addr:#x}")

    def read(self, addr, n):
        buf = C.create_string_buffer(n)
        if self.L.emu_read(self.e, addr, buf, n) != 0:
            raise EmuError(f"read fault at {addr:#x}")
        return buf.raw

    def u64(self, addr):
        return struct.unpack("<Q", self.read(addr, 8))[0]

    def u64s(self, addr, n):
        return list(struct.unpack("<%dQ" % n, self.read(addr, 8 * n)))

    def malloc(self, n):
        p = self.L.emu_malloc(self.e, n)
        if not p:
            raise EmuError("guest heap exhausted")
        return p

    def put(self, data):
        p = self.malloc(max(1, len(data))); self.write(p, data); return p

    # ---- imports implemented in Python ----
    def _unimplemented(self, name):
        def f(rb):
            raise EmuError(f"import {name} is not implemented (called from {rb.m.symbol_at(rb.x[30])})")
        return f

    def imp_write(self):
        fd, p, n = self.x[0], self.x[1], self.x[2]
        self.output.append(self.read(p, n)); self.x[0] = n

    def imp_writev(self):
        fd, iov, cnt = self.x[0], self.x[1], self.x[2]; tot = 0
        for i in range(cnt):
            p, n = self.u64s(iov + 16 * i, 2); self.output.append(self.read(p, n)); tot += n
        self.x[0] = tot

    def imp_getenv(self):
        self.x[0] = 0

    def imp___error(self):
        self.x[0] = self.errno_ptr

    def imp_clock_gettime(self):
        self.write(self.x[1], struct.pack("<qq", 1700000000 + self.icount() // 10**9, self.icount() % 10**9)); self.x[0] = 0

    def imp_CCRandomGenerateBytes(self):
        p, n = self.x[0], self.x[1]
        self.write(p, bytes((0xA5 ^ (i * 37)) & 0xFF for i in range(n))); self.x[0] = 0

    def imp_sysconf(self):
        self.x[0] = 16384 if self.x[0] == 29 else 8

    def imp_pthread_self(self):
        self.x[0] = self.tls

    def imp_pthread_get_stackaddr_np(self):
        self.x[0] = STACK_BASE + STACK_SIZE

    def imp_pthread_get_stacksize_np(self):
        self.x[0] = STACK_SIZE

    def imp_abort(self):
        raise EmuError("guest called abort(): " + self.text_output()[-2000:])

    def imp_exit(self):
        raise EmuError(f"guest called exit({self.x[0]})")

    def imp__Unwind_RaiseException(self):
        raise EmuError("guest panicked (unwinding): " + self.text_output()[-2000:])

    def imp__Unwind_Backtrace(self):
        self.x[0] = 5

    def imp_mmap(self):
        n = self.x[1]; p = self.malloc(n + 32768); p = (p + 16383) & ~16383; self.write(p, b"\0" * min(n, 1 << 20)); self.x[0] = p

    def text_output(self):
        return b"".join(self.output).decode("utf-8", "replace")

    def icount(self):
        return self.L.emu_get(self.e, 3)

    # ---- calling guest functions ----
    def hook(self, addr, fn):
        """fn(rb) runs when the guest reaches addr.  Return None to return to the caller (like a replaced function) or "continue" to execute on."""
        self.L.emu_hook(self.e, addr, 1); self.py_hooks[addr] = fn

    def call(self, addr, args=(), x8=None, stack_args=(), max_instr=10**12, stack_blob=b""):
        """AAPCS64 call: up to 8 integer arguments in x0..x7, indirect-result pointer in x8, further arguments on the stack (8 bytes each)."""
        L, e, x = self.L, self.e, self.x
        if isinstance(addr, str):
            addr = self.m.find(addr, 0)
        for i in range(31):
            x[i] = 0
        for i, a in enumerate(args):
            x[i] = a & 0xFFFFFFFFFFFFFFFF
        if x8 is not None:
            x[8] = x8
        sp = STACK_BASE + STACK_SIZE - 0x10000
        if stack_args or stack_blob:      # stack_blob: pre-packed argument area (Darwin packs sub-8-byte stack arguments to their own size)
            blob = stack_blob or b"".join(struct.pack("<Q", a & 0xFFFFFFFFFFFFFFFF) for a in stack_args)
            sp -= (len(blob) + 15) & ~15; self.write(sp, blob)
        x[30] = STOP_PC; x[29] = 0
        L.emu_set(e, 0, sp); L.emu_set(e, 1, addr)
        while True:
            r = L.emu_run(e, max_instr)
            if r == R_DONE:
                return x[0]
            pc = L.emu_get(e, 1)
            if r == R_HOOK:
                fn = self.py_hooks.get(pc)
                if fn is None:
                    raise EmuError(f"hook without handler at {pc:#x}")
                if fn(self) == "continue":
                    rr = L.emu_step_over(e)
                    if rr != -1:
                        r = rr
                    else:
                        continue
                else:
                    L.emu_return_from_hook(e); continue
            where = self.m.symbol_at(pc)
            if r == R_UNKNOWN:
                raise EmuError(f"undecoded instruction {L.emu_get(e, 5):#010x} at {pc:#x} ({where})")
            if r == R_FAULT:
                raise EmuError(f"memory fault at address {L.emu_get(e, 4):#x}, pc {pc:#x} ({where}), lr {self.m.symbol_at(x[30])}")
            if r == R_TRAP:
                raise EmuError(f"trap (brk) at {pc:#x} ({where}); guest output: {self.text_output()[-1500:]}")
            if r == R_LIMIT:
                raise EmuError(f"instruction limit reached at {pc:#x} ({where})")
            raise EmuError(f"stopped with reason {r}")

    def close(self):
        if self.e:
            self.L.emu_destroy(self.e); self.e = None
