// a64emu.cpp — a small AArch64 user-mode interpreter.  TEST INFRASTRUCTURE (oracle/): it exists for one purpose, to EXECUTE functions of
// the reference's shipped binary `/root/reference/test-dist/xfg-stark-cli` (Mach-O arm64, rustc 1.87, winterfell 0.8.3 statically linked) in
// this x86-64 container, so that the CPU oracle can be pinned against outputs of the reference's own machine code (its Winterfell prover,
// hashes, FFTs, FRI folding, serialisers) instead of against a recollection of the upstream source.  Nothing in the product links or loads it.
//
// Scope: the A64 integer ISA as rustc/LLVM emit it (data processing, loads/stores incl. pairs / exclusives / LSE atomics, branches, system
// hints) plus the AdvSIMD / scalar-FP subset met on the executed paths.  Anything undecoded stops the run with reason UNKNOWN and the word, so a
// gap can never turn into a silently wrong result.  libSystem imports are not emulated as code: the loader (oracle/a64emu/refbin.py) binds each
// `__stubs` entry to a "native" (malloc/free/memcpy/... implemented here on guest memory) or to a Python callback.
//
// Build: g++ -O2 -shared -fPIC -o oracle/_ref/liba64emu.so oracle/a64emu/a64emu.cpp   (oracle/Makefile target `a64emu`)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <vector>

typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64; typedef int64_t s64; typedef int32_t s32;
typedef unsigned __int128 u128;

namespace {

enum Reason : int { R_DONE = 0, R_HOOK = 1, R_UNKNOWN = 2, R_FAULT = 3, R_TRAP = 4, R_LIMIT = 5 };
enum Native : u8 { N_NONE = 0, N_PY = 1, N_MALLOC, N_FREE, N_CALLOC, N_REALLOC, N_MEMALIGN, N_MEMCPY, N_MEMMOVE, N_MEMSET, N_MEMCMP, N_BZERO, N_STRLEN, N_RET0, N_TLV };

struct Region { u64 base, size; u8* host; };

struct VReg { u64 d[2]; };

struct Emu {
  u64 x[32];           // x[31] unused (zr / sp handled by accessors)
  u64 sp, pc;
  u32 n, z, c, v;      // flags
  VReg q[32];
  u64 tpidr;
  u64 icount;
  u64 stop_pc;         // returning here ends the run
  // exit info
  int reason; u64 fault_addr; u32 insn; u64 hook_pc;
  // exclusive monitor
  u64 excl_addr; int excl_valid;
  // memory
  std::vector<Region> regions; Region* last = nullptr;
  // hooks: one byte per instruction word of [text_base, text_base + text_size)
  u64 text_base = 0, text_size = 0; std::vector<u8> flags;
  // heap
  u64 heap_base = 0, heap_size = 0, heap_top = 0; u64 free_list[64] = {0};
  u64 alloc_bytes = 0, alloc_calls = 0;
  u32 fpcr = 0, fpsr = 0;
  u64 tls_base = 0;
};

struct Fault { u64 addr; };

inline u8* xlat(Emu* e, u64 a, u64 n) {
  Region* r = e->last;
  if (r && a - r->base < r->size && a - r->base + n <= r->size) return r->host + (a - r->base);
  for (auto& g : e->regions) if (a - g.base < g.size && a - g.base + n <= g.size) { e->last = &g; return g.host + (a - g.base); }
  throw Fault{a};
}
template <class T> inline T ld(Emu* e, u64 a) { T v; memcpy(&v, xlat(e, a, sizeof(T)), sizeof(T)); return v; }
template <class T> inline void st(Emu* e, u64 a, T v) { memcpy(xlat(e, a, sizeof(T)), &v, sizeof(T)); }

// ---- heap natives (size header 16 bytes before the block; power-of-two size classes) ----
inline int size_class(u64 n) { int k = 5; while ((u64(1) << k) < n) k++; return k; }
u64 g_malloc(Emu* e, u64 n, u64 align = 16) {
  if (align > 16) {   // over-aligned: carve from the bump pointer directly, never recycled
    u64 p = (e->heap_top + 16 + align - 1) & ~(align - 1);
    if (p + n > e->heap_base + e->heap_size) return 0;
    st<u64>(e, p - 16, n); st<u64>(e, p - 8, 0xA11C0D00ull | 63); e->heap_top = p + n; return p;
  }
  const int k = size_class(n + 16);
  e->alloc_calls++; e->alloc_bytes += n;
  if (e->free_list[k]) { u64 p = e->free_list[k]; e->free_list[k] = ld<u64>(e, p); st<u64>(e, p - 16, n); return p; }
  u64 blk = (e->heap_top + 15) & ~u64(15);
  if (blk + (u64(1) << k) > e->heap_base + e->heap_size) return 0;
  e->heap_top = blk + (u64(1) << k);
  st<u64>(e, blk, n); st<u64>(e, blk + 8, 0xA11C0D00ull | k);
  return blk + 16;
}
void g_free(Emu* e, u64 p) {
  if (!p) return;
  const u64 tag = ld<u64>(e, p - 8);
  if ((tag & ~u64(63)) != 0xA11C0D00ull) throw Fault{p};
  const int k = (int)(tag & 63);
  if (k == 63) return;
  st<u64>(e, p, e->free_list[k]); e->free_list[k] = p;
}

// ---- helpers ----
inline u64 ones(int n) { return n >= 64 ? ~u64(0) : ((u64(1) << n) - 1); }
inline u64 ror(u64 v, int r, int size) { r %= size; if (!r) return v & ones(size); v &= ones(size); return ((v >> r) | (v << (size - r))) & ones(size); }
inline int hsb(u32 v) { return 31 - __builtin_clz(v); }
bool decode_bitmasks(int N, int imms, int immr, bool immediate, int datasize, u64& wmask, u64& tmask) {
  const u32 comb = (u32(N) << 6) | (u32(~imms) & 0x3f);
  if (!comb) return false;
  const int len = hsb(comb);
  if (len < 1) return false;
  const int levels = (1 << len) - 1;
  if (immediate && (imms & levels) == levels) return false;
  const int S = imms & levels, R = immr & levels, diff = (S - R) & levels, esize = 1 << len;
  u64 welem = ones(S + 1), telem = ones(diff + 1);
  welem = ror(welem, R, esize);
  wmask = 0; tmask = 0;
  for (int i = 0; i < datasize; i += esize) { wmask |= welem << i; tmask |= telem << i; }
  if (datasize == 32) { wmask &= 0xffffffffull; tmask &= 0xffffffffull; }
  return true;
}
inline u64 sext(u64 v, int bits) { const u64 m = u64(1) << (bits - 1); v &= ones(bits); return (v ^ m) - m; }
inline u32 bits(u32 w, int hi, int lo) { return (w >> lo) & ((1u << (hi - lo + 1)) - 1); }

inline bool cond_holds(Emu* e, u32 cond) {
  bool r;
  switch (cond >> 1) {
    case 0: r = e->z; break;
    case 1: r = e->c; break;
    case 2: r = e->n; break;
    case 3: r = e->v; break;
    case 4: r = e->c && !e->z; break;
    case 5: r = e->n == e->v; break;
    case 6: r = e->n == e->v && !e->z; break;
    default: return true;
  }
  return (cond & 1) ? !r : r;
}
inline u64 add_with_carry(Emu* e, u64 a, u64 b, u32 cin, bool sf, bool setflags) {
  if (sf) {
    const u128 us = (u128)a + b + cin; const u64 r = (u64)us;
    if (setflags) { e->n = r >> 63; e->z = r == 0; e->c = (u32)(us >> 64); e->v = ((~(a ^ b) & (a ^ r)) >> 63) & 1; }
    return r;
  }
  const u32 a32 = (u32)a, b32 = (u32)b; const u64 us = (u64)a32 + b32 + cin; const u32 r = (u32)us;
  if (setflags) { e->n = r >> 31; e->z = r == 0; e->c = (u32)(us >> 32); e->v = ((~(a32 ^ b32) & (a32 ^ r)) >> 31) & 1; }
  return r;
}
inline u64 shift_reg(u64 v, int type, int amt, bool sf) {
  const int size = sf ? 64 : 32; if (!sf) v &= 0xffffffffull;
  if (!amt) return v;
  switch (type) {
    case 0: return sf ? v << amt : (u32)((u32)v << amt);
    case 1: return v >> amt;
    case 2: return sf ? (u64)((s64)v >> amt) : (u32)((s32)(u32)v >> amt);
    default: return ror(v, amt, size);
  }
}
inline u64 extend_reg(u64 v, int option, int shift) {
  switch (option) {
    case 0: v = (u8)v; break; case 1: v = (u16)v; break; case 2: v = (u32)v; break; case 3: break;
    case 4: v = (u64)(s64)(int8_t)v; break; case 5: v = (u64)(s64)(int16_t)v; break; case 6: v = (u64)(s64)(s32)v; break; default: break;
  }
  return v << shift;
}

#define XR(n) ((n) == 31 ? u64(0) : e->x[n])
#define XSP(n) ((n) == 31 ? e->sp : e->x[n])
#define SETX(n, val) do { if ((n) != 31) e->x[n] = (val); } while (0)
#define SETXSP(n, val) do { if ((n) == 31) e->sp = (val); else e->x[n] = (val); } while (0)
#define UNK() do { e->reason = R_UNKNOWN; e->insn = w; return false; } while (0)

// element accessors of a 128-bit register
template <class T> inline T vget(const VReg& r, int i) { T v; memcpy(&v, (const u8*)r.d + i * sizeof(T), sizeof(T)); return v; }
template <class T> inline void vset(VReg& r, int i, T v) { memcpy((u8*)r.d + i * sizeof(T), &v, sizeof(T)); }
inline u64 velem(const VReg& r, int esz_log, int i) {
  switch (esz_log) { case 0: return vget<u8>(r, i); case 1: return vget<u16>(r, i); case 2: return vget<u32>(r, i); default: return vget<u64>(r, i); }
}
inline void vsetelem(VReg& r, int esz_log, int i, u64 v) {
  switch (esz_log) { case 0: vset<u8>(r, i, (u8)v); break; case 1: vset<u16>(r, i, (u16)v); break; case 2: vset<u32>(r, i, (u32)v); break; default: vset<u64>(r, i, v); }
}

u64 advsimd_expand_imm(int op, int cmode, u32 imm8) {
  u64 imm = 0;
  switch (cmode >> 1) {
    case 0: imm = imm8; imm |= imm << 32; break;
    case 1: imm = (u64)imm8 << 8; imm |= imm << 32; break;
    case 2: imm = (u64)imm8 << 16; imm |= imm << 32; break;
    case 3: imm = (u64)imm8 << 24; imm |= imm << 32; break;
    case 4: imm = imm8; imm |= imm << 16; imm |= imm << 32; break;
    case 5: imm = (u64)imm8 << 8; imm |= imm << 16; imm |= imm << 32; break;
    case 6: imm = (cmode & 1) ? (((u64)imm8 << 16) | 0xffff) : (((u64)imm8 << 8) | 0xff); imm |= imm << 32; break;
    case 7:
      if (!(cmode & 1) && !op) { imm = imm8; imm |= imm << 8; imm |= imm << 16; imm |= imm << 32; }
      else if (!(cmode & 1) && op) { for (int i = 0; i < 8; i++) if (imm8 & (1u << i)) imm |= u64(0xff) << (8 * i); }
      else if ((cmode & 1) && !op) {   // fp32 immediate
        const u32 a = imm8 >> 7, b = (imm8 >> 6) & 1, cdefgh = imm8 & 0x3f;
        u32 f = (a << 31) | ((b ? 0x1fu : 0x20u) << 25) | (cdefgh << 19); imm = f; imm |= imm << 32;
      } else {                          // fp64 immediate
        const u64 a = imm8 >> 7, b = (imm8 >> 6) & 1, cdefgh = imm8 & 0x3f;
        imm = (a << 63) | ((b ? u64(0xff) : u64(0x100)) << 54) | (cdefgh << 48);
      }
      break;
  }
  return imm;
}

bool exec_simd(Emu* e, u32 w);
bool exec_ldst(Emu* e, u32 w);

// executes one instruction at e->pc (pc is advanced / branched here).  false = stop (reason set)
bool step(Emu* e) {
  const u32 w = ld<u32>(e, e->pc);
  const u32 op0 = (w >> 25) & 0xf;
  u64 next = e->pc + 4;
  switch (op0) {
    case 8: case 9: {   // data processing - immediate
      const u32 op = (w >> 23) & 7; const bool sf = w >> 31; const u32 rd = w & 31, rn = (w >> 5) & 31;
      if (op == 0 || op == 1) {   // ADR / ADRP
        const s64 imm = (s64)sext(((u64)bits(w, 23, 5) << 2) | bits(w, 30, 29), 21);
        SETX(rd, (w >> 31) ? ((e->pc & ~u64(0xfff)) + (u64)(imm << 12)) : (e->pc + (u64)imm));
      } else if (op == 2) {       // ADD/SUB immediate
        u64 imm = bits(w, 21, 10); if (w & (1u << 22)) imm <<= 12;
        const bool sub = (w >> 30) & 1, S = (w >> 29) & 1;
        const u64 a = XSP(rn);
        const u64 r = sub ? add_with_carry(e, a, ~imm, 1, sf, S) : add_with_carry(e, a, imm, 0, sf, S);
        if (S) SETX(rd, r); else SETXSP(rd, sf ? r : (u32)r);
      } else if (op == 4) {       // logical immediate
        u64 wm, tm; if (!decode_bitmasks((w >> 22) & 1, bits(w, 15, 10), bits(w, 21, 16), true, sf ? 64 : 32, wm, tm)) UNK();
        const u32 opc = bits(w, 30, 29); u64 a = XR(rn), r;
        switch (opc) { case 0: r = a & wm; break; case 1: r = a | wm; break; case 2: r = a ^ wm; break; default: r = a & wm; }
        if (!sf) r = (u32)r;
        if (opc == 3) { e->n = sf ? r >> 63 : (r >> 31) & 1; e->z = r == 0; e->c = 0; e->v = 0; SETX(rd, r); } else SETXSP(rd, r);
      } else if (op == 5) {       // move wide
        const u32 opc = bits(w, 30, 29), hw = bits(w, 22, 21); const u64 imm = (u64)bits(w, 20, 5) << (16 * hw);
        if (opc == 0) { u64 r = ~imm; SETX(rd, sf ? r : (u32)r); }
        else if (opc == 2) SETX(rd, imm);
        else if (opc == 3) { u64 r = (XR(rd) & ~(u64(0xffff) << (16 * hw))) | imm; SETX(rd, sf ? r : (u32)r); }
        else UNK();
      } else if (op == 6) {       // bitfield
        const u32 opc = bits(w, 30, 29); const int immr = bits(w, 21, 16), imms = bits(w, 15, 10), ds = sf ? 64 : 32;
        u64 wm, tm; if (!decode_bitmasks((w >> 22) & 1, imms, immr, false, ds, wm, tm)) UNK();
        const u64 src = XR(rn) & ones(ds), dst = opc == 1 ? (XR(rd) & ones(ds)) : 0;
        const u64 bot = (dst & ~wm) | (ror(src, immr, ds) & wm);
        u64 top = dst; if (opc == 0) top = ((src >> imms) & 1) ? ones(ds) : 0;
        if (opc == 3) UNK();
        const u64 r = ((top & ~tm) | (bot & tm)) & ones(ds);
        SETX(rd, r);
      } else if (op == 7) {       // EXTR
        const u32 rm = (w >> 16) & 31; const int lsb = bits(w, 15, 10), ds = sf ? 64 : 32;
        const u64 hi = XR(rn) & ones(ds), lo = XR(rm) & ones(ds);
        u64 r = lsb ? ((lo >> lsb) | (hi << (ds - lsb))) : lo;
        SETX(rd, r & ones(ds));
      } else UNK();
      break;
    }
    case 10: case 11: {  // branches, exception generation, system
      if ((w & 0x7c000000) == 0x14000000) {   // B / BL
        const u64 tgt = e->pc + (u64)((s64)sext(w & 0x3ffffff, 26) << 2);
        if (w >> 31) e->x[30] = e->pc + 4;
        next = tgt;
      } else if ((w & 0x7e000000) == 0x34000000) {   // CBZ / CBNZ
        const bool sf = w >> 31; u64 v = XR(w & 31); if (!sf) v = (u32)v;
        if ((v == 0) != (bool)((w >> 24) & 1)) next = e->pc + (u64)((s64)sext(bits(w, 23, 5), 19) << 2);
      } else if ((w & 0x7e000000) == 0x36000000) {   // TBZ / TBNZ
        const int bit = (int)(((w >> 31) << 5) | bits(w, 23, 19));
        const bool set = (XR(w & 31) >> bit) & 1;
        if (set == (bool)((w >> 24) & 1)) next = e->pc + (u64)((s64)sext(bits(w, 18, 5), 14) << 2);
      } else if ((w & 0xff000010) == 0x54000000) {   // B.cond
        if (cond_holds(e, w & 15)) next = e->pc + (u64)((s64)sext(bits(w, 23, 5), 19) << 2);
      } else if ((w & 0xfe1ffc1f) == 0xd61f0000) {   // BR / BLR / RET
        const u32 opc = bits(w, 24, 21); const u64 tgt = XR((w >> 5) & 31);
        if (opc == 1) e->x[30] = e->pc + 4; else if (opc != 0 && opc != 2) UNK();
        next = tgt;
      } else if ((w & 0xfffff01f) == 0xd503201f) {   // hints (NOP, YIELD, ...)
      } else if ((w & 0xfffff01f) == 0xd503301f) {   // barriers (DSB/DMB/ISB/CLREX)
        if (bits(w, 7, 5) == 2) e->excl_valid = 0;
      } else if ((w & 0xfff00000) == 0xd5300000 || (w & 0xfff00000) == 0xd5100000) {   // MRS / MSR (register)
        const bool rd = (w >> 21) & 1; const u32 sys = bits(w, 19, 5), rt = w & 31;
        // op0:op1:CRn:CRm:op2 as 15 bits (o0 is bit 19 -> op0 = 2 + o0)
        if (sys == 0x5e82 || sys == 0x5e83) { if (rd) SETX(rt, e->tpidr); else e->tpidr = XR(rt); }      // TPIDR_EL0 / TPIDRRO_EL0
        else if (sys == 0x5a20) { if (rd) SETX(rt, e->fpcr); else e->fpcr = (u32)XR(rt); }               // FPCR
        else if (sys == 0x5a21) { if (rd) SETX(rt, e->fpsr); else e->fpsr = (u32)XR(rt); }               // FPSR
        else if (sys == 0x5807 && rd) SETX(rt, 0x8444c004);                                             // CTR_EL0
        else if (sys == 0x5f02 && rd) SETX(rt, e->icount);                                              // CNTVCT_EL0
        else if (sys == 0x5f00 && rd) SETX(rt, 24000000);                                               // CNTFRQ_EL0
        else if (sys == 0x5807) {} else UNK();
      } else if ((w & 0xffe0001f) == 0xd4200000) {   // BRK
        e->reason = R_TRAP; e->insn = w; return false;
      } else UNK();
      break;
    }
    case 4: case 6: case 12: case 14:   // loads and stores
      if (!exec_ldst(e, w)) return false;
      break;
    case 5: case 13: {   // data processing - register
      const bool sf = w >> 31; const u32 rd = w & 31, rn = (w >> 5) & 31, rm = (w >> 16) & 31;
      if (!((w >> 28) & 1)) {
        if (!((w >> 24) & 1)) {   // logical shifted register
          const u32 opc = bits(w, 30, 29); const bool N = (w >> 21) & 1;
          u64 b = shift_reg(XR(rm), bits(w, 23, 22), bits(w, 15, 10), sf); if (N) b = ~b;
          u64 a = XR(rn), r;
          switch (opc) { case 0: r = a & b; break; case 1: r = a | b; break; case 2: r = a ^ b; break; default: r = a & b; }
          if (!sf) r = (u32)r;
          if (opc == 3) { e->n = sf ? r >> 63 : (r >> 31) & 1; e->z = r == 0; e->c = 0; e->v = 0; }
          SETX(rd, r);
        } else if (!((w >> 21) & 1)) {   // add/sub shifted register
          const bool sub = (w >> 30) & 1, S = (w >> 29) & 1;
          const u64 b = shift_reg(XR(rm), bits(w, 23, 22), bits(w, 15, 10), sf), a = XR(rn);
          const u64 r = sub ? add_with_carry(e, a, ~b, 1, sf, S) : add_with_carry(e, a, b, 0, sf, S);
          SETX(rd, sf ? r : (u32)r);
        } else {                          // add/sub extended register
          const bool sub = (w >> 30) & 1, S = (w >> 29) & 1;
          const u64 b = extend_reg(XR(rm), bits(w, 15, 13), bits(w, 12, 10)), a = XSP(rn);
          const u64 r = sub ? add_with_carry(e, a, ~b, 1, sf, S) : add_with_carry(e, a, b, 0, sf, S);
          if (S) SETX(rd, sf ? r : (u32)r); else SETXSP(rd, sf ? r : (u32)r);
        }
      } else if ((w >> 24) & 1) {   // data processing 3 source
        const u32 op31 = bits(w, 23, 21), ra = bits(w, 14, 10); const bool o0 = (w >> 15) & 1;
        u64 r;
        if (op31 == 0) { const u64 p = XR(rn) * XR(rm); r = o0 ? XR(ra) - p : XR(ra) + p; if (!sf) r = (u32)r; }
        else if (op31 == 1) { const s64 p = (s64)(s32)XR(rn) * (s64)(s32)XR(rm); r = o0 ? XR(ra) - (u64)p : XR(ra) + (u64)p; }
        else if (op31 == 5) { const u64 p = (u64)(u32)XR(rn) * (u64)(u32)XR(rm); r = o0 ? XR(ra) - p : XR(ra) + p; }
        else if (op31 == 2) r = (u64)(((__int128)(s64)XR(rn) * (__int128)(s64)XR(rm)) >> 64);
        else if (op31 == 6) r = (u64)(((u128)XR(rn) * (u128)XR(rm)) >> 64);
        else UNK();
        SETX(rd, r);
      } else {
        const u32 grp = bits(w, 23, 21);
        if (grp == 0) {          // ADC / SBC
          if (bits(w, 15, 10)) UNK();
          const bool sub = (w >> 30) & 1, S = (w >> 29) & 1; u64 b = XR(rm); if (sub) b = ~b; if (!sf) b = (u32)b;
          const u64 r = add_with_carry(e, XR(rn), b, e->c, sf, S); SETX(rd, sf ? r : (u32)r);
        } else if (grp == 2) {   // conditional compare
          const bool sub = (w >> 30) & 1; const u32 cond = bits(w, 15, 12), nzcv = w & 15;
          u64 b = ((w >> 11) & 1) ? (u64)rm : XR(rm);
          if (cond_holds(e, cond)) { if (sub) add_with_carry(e, XR(rn), sf ? ~b : (u32)~b, 1, sf, true); else add_with_carry(e, XR(rn), sf ? b : (u32)b, 0, sf, true); }
          else { e->n = (nzcv >> 3) & 1; e->z = (nzcv >> 2) & 1; e->c = (nzcv >> 1) & 1; e->v = nzcv & 1; }
        } else if (grp == 4) {   // conditional select
          const bool op = (w >> 30) & 1; const u32 op2 = bits(w, 11, 10), cond = bits(w, 15, 12);
          u64 r;
          if (cond_holds(e, cond)) r = XR(rn);
          else { r = XR(rm); if (op) r = ~r; if (op2 & 1) r += 1; }
          if (op2 > 1) UNK();
          SETX(rd, sf ? r : (u32)r);
        } else if (grp == 6) {   // data processing 1 / 2 source
          const u32 opc = bits(w, 15, 10);
          if ((w >> 30) & 1) {   // 1 source
            u64 a = XR(rn), r; const int ds = sf ? 64 : 32; if (!sf) a = (u32)a;
            switch (opc) {
              case 0: { r = 0; for (int i = 0; i < ds; i++) if ((a >> i) & 1) r |= u64(1) << (ds - 1 - i); break; }                 // RBIT
              case 1: { r = 0; for (int i = 0; i < ds / 8; i++) r |= ((a >> (8 * i)) & 0xff) << (8 * (i ^ 1)); break; }              // REV16
              case 2: if (sf) { r = ((u64)__builtin_bswap32((u32)(a >> 32)) << 32) | __builtin_bswap32((u32)a); } else r = __builtin_bswap32((u32)a); break;
              case 3: if (!sf) UNK(); r = __builtin_bswap64(a); break;
              case 4: r = a ? (sf ? __builtin_clzll(a) : __builtin_clz((u32)a)) : ds; break;                                            // CLZ
              case 5: { const u64 sgn = (a >> (ds - 1)) & 1; int k = 0; for (int i = ds - 2; i >= 0 && (((a >> i) & 1) == sgn); i--) k++; r = k; break; }  // CLS
              default: UNK();
            }
            SETX(rd, r);
          } else {
            u64 a = XR(rn), b = XR(rm), r; const int ds = sf ? 64 : 32; if (!sf) { a = (u32)a; b = (u32)b; }
            switch (opc) {
              case 2: r = b ? a / b : 0; break;                                                                                         // UDIV
              case 3: if (sf) { const s64 sa = (s64)a, sb = (s64)b; r = sb ? (sa == INT64_MIN && sb == -1 ? (u64)sa : (u64)(sa / sb)) : 0; }
                      else { const s32 sa = (s32)a, sb = (s32)b; r = sb ? (sa == INT32_MIN && sb == -1 ? (u32)sa : (u32)(sa / sb)) : 0; } break;
              case 8: r = shift_reg(a, 0, (int)(b % ds), sf); break;
              case 9: r = shift_reg(a, 1, (int)(b % ds), sf); break;
              case 10: r = shift_reg(a, 2, (int)(b % ds), sf); break;
              case 11: r = shift_reg(a, 3, (int)(b % ds), sf); break;
              default: UNK();
            }
            SETX(rd, sf ? r : (u32)r);
          }
        } else UNK();
      }
      break;
    }
    case 7: case 15:
      if (!exec_simd(e, w)) return false;
      break;
    default: UNK();
  }
  e->pc = next;
  return true;
}

// ---- loads and stores ----
inline void vload(Emu* e, VReg& r, u64 a, int bytes) { r.d[0] = r.d[1] = 0; memcpy(r.d, xlat(e, a, bytes), bytes); }
inline void vstore(Emu* e, const VReg& r, u64 a, int bytes) { memcpy(xlat(e, a, bytes), r.d, bytes); }

bool exec_ldst(Emu* e, u32 w) {
  const u32 rt = w & 31, rn = (w >> 5) & 31;
  const u32 top = bits(w, 29, 27);
  const bool V = (w >> 26) & 1;
  if (top == 1 && !V && bits(w, 25, 24) == 0) {   // exclusive / ordered / CAS:  size 001000 o2 L o1 Rs o0 Rt2 Rn Rt
    const u32 size = w >> 30, o2 = (w >> 23) & 1, L = (w >> 22) & 1, o1 = (w >> 21) & 1, rs = (w >> 16) & 31, rt2 = bits(w, 14, 10);
    const u64 a = XSP(rn); const int bytes = 1 << size;
    auto load = [&](u64 addr) -> u64 { switch (size) { case 0: return ld<u8>(e, addr); case 1: return ld<u16>(e, addr); case 2: return ld<u32>(e, addr); default: return ld<u64>(e, addr); } };
    auto store = [&](u64 addr, u64 v) { switch (size) { case 0: st<u8>(e, addr, (u8)v); break; case 1: st<u16>(e, addr, (u16)v); break; case 2: st<u32>(e, addr, (u32)v); break; default: st<u64>(e, addr, v); } };
    if (o2 && o1) {          // CAS / CASA / CASL / CASAL (rt2 must be 31)
      if (rt2 != 31) UNK();
      const u64 cur = load(a), cmp = XR(rs) & ones(8 * bytes);
      if (cur == cmp) store(a, XR(rt));
      SETX(rs, cur);
    } else if (!o2 && o1) {  // LDXP / STXP (pair)
      if (L) { if (size == 3) { SETX(rt, ld<u64>(e, a)); SETX(rt2, ld<u64>(e, a + 8)); } else { SETX(rt, ld<u32>(e, a)); SETX(rt2, ld<u32>(e, a + 4)); } e->excl_addr = a; e->excl_valid = 1; }
      else { if (size == 3) { st<u64>(e, a, XR(rt)); st<u64>(e, a + 8, XR(rt2)); } else { st<u32>(e, a, (u32)XR(rt)); st<u32>(e, a + 4, (u32)XR(rt2)); } SETX(rs, 0); e->excl_valid = 0; }
    } else if (!o2) {        // LDXR / LDAXR / STXR / STLXR
      if (L) { SETX(rt, load(a)); e->excl_addr = a; e->excl_valid = 1; }
      else { store(a, XR(rt)); SETX(rs, 0); e->excl_valid = 0; }   // single-threaded: always succeeds
    } else {                 // LDAR / STLR (and LDLAR / STLLR)
      if (L) SETX(rt, load(a)); else store(a, XR(rt));
    }
    return true;
  }
  if (top == 3 && bits(w, 25, 24) == 0) {   // load register (literal)
    const u32 opc = w >> 30; const u64 a = e->pc + (u64)((s64)sext(bits(w, 23, 5), 19) << 2);
    if (V) { vload(e, e->q[rt], a, 4 << opc); if (opc > 2) UNK(); }
    else if (opc == 0) SETX(rt, ld<u32>(e, a)); else if (opc == 1) SETX(rt, ld<u64>(e, a)); else if (opc == 2) SETX(rt, (u64)(s64)(s32)ld<u32>(e, a)); else {}
    return true;
  }
  if (top == 5) {   // load/store pair
    const u32 opc = w >> 30, type = bits(w, 24, 23), L = (w >> 22) & 1, rt2 = bits(w, 14, 10);
    int scale; if (V) scale = 2 + opc; else scale = (opc & 2) ? 3 : 2;
    if (V && opc == 3) UNK();
    const s64 off = (s64)sext(bits(w, 21, 15), 7) << scale;
    u64 base = XSP(rn); const u64 a = (type == 1) ? base : base + (u64)off;
    const int bytes = 1 << scale;
    if (V) {
      if (L) { vload(e, e->q[rt], a, bytes); vload(e, e->q[rt2], a + bytes, bytes); }
      else { vstore(e, e->q[rt], a, bytes); vstore(e, e->q[rt2], a + bytes, bytes); }
    } else if (L) {
      if (opc == 1) { const u64 v1 = (u64)(s64)(s32)ld<u32>(e, a), v2 = (u64)(s64)(s32)ld<u32>(e, a + 4); SETX(rt, v1); SETX(rt2, v2); }
      else if (scale == 3) { const u64 v1 = ld<u64>(e, a), v2 = ld<u64>(e, a + 8); SETX(rt, v1); SETX(rt2, v2); }
      else { const u64 v1 = ld<u32>(e, a), v2 = ld<u32>(e, a + 4); SETX(rt, v1); SETX(rt2, v2); }
    } else {
      if (scale == 3) { st<u64>(e, a, XR(rt)); st<u64>(e, a + 8, XR(rt2)); } else { st<u32>(e, a, (u32)XR(rt)); st<u32>(e, a + 4, (u32)XR(rt2)); }
    }
    if (type == 1 || type == 3) SETXSP(rn, base + (u64)off);
    return true;
  }
  if (top == 7) {   // load/store register
    const u32 size = w >> 30, opc = bits(w, 23, 22);
    u64 a; bool wb = false; u64 wbv = 0;
    if ((w >> 24) & 1) {   // unsigned offset
      int scale = size; if (V && (opc & 2)) scale = 4;
      a = XSP(rn) + ((u64)bits(w, 21, 10) << scale);
    } else if (!((w >> 21) & 1)) {
      const s64 imm = (s64)sext(bits(w, 20, 12), 9); const u32 mode = bits(w, 11, 10);
      const u64 base = XSP(rn);
      if (mode == 0 || mode == 2) a = base + (u64)imm; else if (mode == 1) { a = base; wb = true; wbv = base + (u64)imm; } else { a = base + (u64)imm; wb = true; wbv = a; }
    } else if (bits(w, 11, 10) == 2) {   // register offset
      int scale = size; if (V && (opc & 2)) scale = 4;
      const u32 option = bits(w, 15, 13); const int sh = ((w >> 12) & 1) ? scale : 0;
      a = XSP(rn) + extend_reg(XR((w >> 16) & 31), option, sh);
    } else if (bits(w, 11, 10) == 0 && !V) {   // LSE atomic memory operations
      const u32 rs = (w >> 16) & 31, o3 = (w >> 15) & 1, op = bits(w, 14, 12);
      const u64 addr = XSP(rn); const int nb = 8 << size;
      u64 old; switch (size) { case 0: old = ld<u8>(e, addr); break; case 1: old = ld<u16>(e, addr); break; case 2: old = ld<u32>(e, addr); break; default: old = ld<u64>(e, addr); }
      const u64 val = XR(rs) & ones(nb); u64 nv;
      if (o3 && op == 4 && rs == 31 && bits(w, 23, 22) == 2) { SETX(rt, old); return true; }   // LDAPR
      if (o3) { if (op != 0) UNK(); nv = val; }
      else switch (op) {
        case 0: nv = old + val; break; case 1: nv = old & ~val; break; case 2: nv = old ^ val; break; case 3: nv = old | val; break;
        case 4: nv = (s64)sext(old, nb) > (s64)sext(val, nb) ? old : val; break; case 5: nv = (s64)sext(old, nb) < (s64)sext(val, nb) ? old : val; break;
        case 6: nv = old > val ? old : val; break; default: nv = old < val ? old : val; break;
      }
      nv &= ones(nb);
      switch (size) { case 0: st<u8>(e, addr, (u8)nv); break; case 1: st<u16>(e, addr, (u16)nv); break; case 2: st<u32>(e, addr, (u32)nv); break; default: st<u64>(e, addr, nv); }
      SETX(rt, old);
      return true;
    } else UNK();
    if (V) {
      int bytes = 1 << size; if (opc & 2) { if (size != 0) UNK(); bytes = 16; }
      if (opc & 1) vload(e, e->q[rt], a, bytes); else vstore(e, e->q[rt], a, bytes);
    } else {
      if (opc == 0) { switch (size) { case 0: st<u8>(e, a, (u8)XR(rt)); break; case 1: st<u16>(e, a, (u16)XR(rt)); break; case 2: st<u32>(e, a, (u32)XR(rt)); break; default: st<u64>(e, a, XR(rt)); } }
      else if (opc == 1) { u64 v; switch (size) { case 0: v = ld<u8>(e, a); break; case 1: v = ld<u16>(e, a); break; case 2: v = ld<u32>(e, a); break; default: v = ld<u64>(e, a); } SETX(rt, v); }
      else if (size == 3 && opc == 2) { /* PRFM */ }
      else { u64 v; switch (size) { case 0: v = (u64)(s64)(int8_t)ld<u8>(e, a); break; case 1: v = (u64)(s64)(int16_t)ld<u16>(e, a); break; case 2: v = (u64)(s64)(s32)ld<u32>(e, a); break; default: UNK(); }
             if (opc == 3) v = (u32)v; SETX(rt, v); }
    }
    if (wb) SETXSP(rn, wbv);
    return true;
  }
  if (top == 1 && V) {   // AdvSIMD load/store structures
    const bool Q = (w >> 30) & 1, L = (w >> 22) & 1, post = (w >> 23) & 1; const u32 rm = (w >> 16) & 31;
    if (bits(w, 29, 24) == 0x0c) {   // multiple structures
      const u32 opcode = bits(w, 15, 12), size = bits(w, 11, 10);
      int regs; int interleave = 1;
      switch (opcode) { case 7: regs = 1; break; case 10: regs = 2; break; case 6: regs = 3; break; case 2: regs = 4; break;
                        case 8: regs = 2; interleave = 2; break; case 4: regs = 3; interleave = 3; break; case 0: regs = 4; interleave = 4; break; default: UNK(); }
      const int bytes = Q ? 16 : 8; u64 a = XSP(rn); const u64 base = a;
      if (interleave == 1) {
        for (int r = 0; r < regs; r++) { VReg& q = e->q[(rt + r) & 31]; if (L) vload(e, q, a, bytes); else vstore(e, q, a, bytes); a += bytes; }
      } else {
        const int esz = 1 << size, elems = bytes / esz;
        if (L) for (int r = 0; r < regs; r++) e->q[(rt + r) & 31].d[0] = e->q[(rt + r) & 31].d[1] = 0;
        for (int i = 0; i < elems; i++) for (int r = 0; r < regs; r++) {
          VReg& q = e->q[(rt + r) & 31];
          if (L) memcpy((u8*)q.d + i * esz, xlat(e, a, esz), esz); else memcpy(xlat(e, a, esz), (u8*)q.d + i * esz, esz);
          a += esz;
        }
      }
      if (post) SETXSP(rn, rm == 31 ? base + (u64)(bytes * regs) : base + XR(rm));
      return true;
    }
    if (bits(w, 29, 24) == 0x0d) {   // single structure (only the 1-register forms and LD1R)
      const u32 R = (w >> 21) & 1, opcode = bits(w, 15, 13), S = (w >> 12) & 1, size = bits(w, 11, 10);
      if (R) UNK();
      u64 a = XSP(rn); int esz_log, index;
      if (opcode == 6) {   // LD1R
        if (!L || S) UNK();
        esz_log = size; const int esz = 1 << esz_log; u64 v = 0; memcpy(&v, xlat(e, a, esz), esz);
        VReg& q = e->q[rt]; q.d[0] = q.d[1] = 0; for (int i = 0; i < (Q ? 16 : 8) / esz; i++) vsetelem(q, esz_log, i, v);
        if (post) SETXSP(rn, rm == 31 ? a + esz : a + XR(rm));
        return true;
      }
      if (opcode == 0) { esz_log = 0; index = (Q << 3) | (S << 2) | size; }
      else if (opcode == 2) { esz_log = 1; index = (Q << 2) | (S << 1) | (size >> 1); }
      else if (opcode == 4 && !(size & 1)) { esz_log = 2; index = (Q << 1) | S; }
      else if (opcode == 4 && size == 1 && !S) { esz_log = 3; index = Q; }
      else UNK();
      const int esz = 1 << esz_log;
      if (L) { u64 v = 0; memcpy(&v, xlat(e, a, esz), esz); vsetelem(e->q[rt], esz_log, index, v); }
      else { const u64 v = velem(e->q[rt], esz_log, index); memcpy(xlat(e, a, esz), &v, esz); }
      if (post) SETXSP(rn, rm == 31 ? a + esz : a + XR(rm));
      return true;
    }
  }
  UNK();
}

#include "a64simd.inc"

bool do_native(Emu* e, u8 id) {
  u64* x = e->x;
  switch (id) {
    case N_MALLOC: x[0] = g_malloc(e, x[0] ? x[0] : 1); break;
    case N_FREE: g_free(e, x[0]); break;
    case N_CALLOC: { const u64 n = x[0] * x[1]; const u64 p = g_malloc(e, n ? n : 1); if (p) memset(xlat(e, p, n ? n : 1), 0, n ? n : 1); x[0] = p; break; }
    case N_REALLOC: {
      const u64 p = x[0], n = x[1] ? x[1] : 1;
      if (!p) { x[0] = g_malloc(e, n); break; }
      const u64 old = ld<u64>(e, p - 16); const u64 tag = ld<u64>(e, p - 8); const int k = (int)(tag & 63);
      if (k != 63 && n + 16 <= (u64(1) << k)) { st<u64>(e, p - 16, n); break; }
      const u64 q = g_malloc(e, n); if (q) { memcpy(xlat(e, q, n), xlat(e, p, old < n ? old : n), old < n ? old : n); g_free(e, p); } x[0] = q; break;
    }
    case N_MEMALIGN: { const u64 p = g_malloc(e, x[2] ? x[2] : 1, x[1] < 16 ? 16 : x[1]); if (!p) { x[0] = 12; break; } st<u64>(e, x[0], p); x[0] = 0; break; }
    case N_MEMCPY: case N_MEMMOVE: if (x[2]) memmove(xlat(e, x[0], x[2]), xlat(e, x[1], x[2]), x[2]); break;
    case N_MEMSET: if (x[2]) memset(xlat(e, x[0], x[2]), (int)x[1], x[2]); break;
    case N_BZERO: if (x[1]) memset(xlat(e, x[0], x[1]), 0, x[1]); break;
    case N_MEMCMP: x[0] = x[2] ? (u64)(s64)memcmp(xlat(e, x[0], x[2]), xlat(e, x[1], x[2]), x[2]) : 0; break;
    case N_STRLEN: { u64 n = 0; while (ld<u8>(e, x[0] + n)) n++; x[0] = n; break; }
    case N_RET0: x[0] = 0; break;
    case N_TLV: x[0] = e->tls_base + ld<u64>(e, x[0] + 16); break;   // macOS thread-local thunk: x0 = &descriptor {thunk, key, offset} -> address of the variable
    default: return false;
  }
  e->pc = e->x[30];
  return true;
}

}  // namespace

extern "C" {

Emu* emu_create() { Emu* e = new Emu(); memset(e->x, 0, sizeof e->x); e->sp = e->pc = 0; e->n = e->z = e->c = e->v = 0; memset(e->q, 0, sizeof e->q); e->tpidr = 0; e->icount = 0;
  e->stop_pc = 0xDEAD0000ull; e->reason = 0; e->excl_valid = 0; e->regions.reserve(16); return e; }
void emu_destroy(Emu* e) { if (!e) return; for (auto& r : e->regions) munmap(r.host, r.size); delete e; }
int emu_map(Emu* e, u64 base, u64 size) {
  if (e->regions.size() >= 16) return -1;
  void* p = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (p == MAP_FAILED) return -1;
  e->regions.push_back(Region{base, size, (u8*)p}); e->last = nullptr; return 0;
}
int emu_set_heap(Emu* e, u64 base, u64 size) { if (emu_map(e, base, size)) return -1; e->heap_base = base; e->heap_size = size; e->heap_top = base + 64; return 0; }
int emu_write(Emu* e, u64 addr, const void* src, u64 n) { try { if (n) memcpy(xlat(e, addr, n), src, n); return 0; } catch (Fault&) { return -1; } }
int emu_read(Emu* e, u64 addr, void* dst, u64 n) { try { if (n) memcpy(dst, xlat(e, addr, n), n); return 0; } catch (Fault&) { return -1; } }
u64 emu_malloc(Emu* e, u64 n) { try { return g_malloc(e, n ? n : 1); } catch (Fault&) { return 0; } }
void emu_set_text(Emu* e, u64 base, u64 size) { e->text_base = base; e->text_size = size; e->flags.assign(size / 4, 0); }
int emu_hook(Emu* e, u64 addr, int native_id) { if (addr - e->text_base >= e->text_size) return -1; e->flags[(addr - e->text_base) >> 2] = (u8)native_id; return 0; }
u64* emu_regs(Emu* e) { return e->x; }
u64 emu_get(Emu* e, int what) { switch (what) { case 0: return e->sp; case 1: return e->pc; case 2: return (e->n << 3) | (e->z << 2) | (e->c << 1) | e->v; case 3: return e->icount; case 4: return e->fault_addr;
  case 5: return e->insn; case 6: return e->tpidr; case 7: return e->heap_top - e->heap_base; case 8: return e->alloc_calls; default: return 0; } }
void emu_set(Emu* e, int what, u64 v) { switch (what) { case 0: e->sp = v; break; case 1: e->pc = v; break; case 6: e->tpidr = v; break; case 9: e->stop_pc = v; break; case 10: e->tls_base = v; break; default: break; } }
u64* emu_vreg(Emu* e, int i) { return e->q[i & 31].d; }

// runs until the stop address, a Python hook, an undecoded instruction, a fault, a trap or the instruction limit
int emu_run(Emu* e, u64 max_instr) {
  const u64 tb = e->text_base, ts = e->text_size; const u8* fl = e->flags.data();
  const u64 limit = e->icount + max_instr;
  bool skip_hook = e->reason == R_HOOK && e->hook_pc == e->pc && false;
  (void)skip_hook;
  try {
    for (;;) {
      const u64 pc = e->pc;
      if (pc == e->stop_pc) { e->reason = R_DONE; return R_DONE; }
      if (pc - tb < ts) {
        const u8 f = fl[(pc - tb) >> 2];
        if (f) {
          if (f == N_PY) { e->reason = R_HOOK; e->hook_pc = pc; return R_HOOK; }
          if (do_native(e, f)) { e->icount++; continue; }
        }
      }
      if (!step(e)) return e->reason;
      if (++e->icount >= limit) { e->reason = R_LIMIT; return R_LIMIT; }
    }
  } catch (Fault& f) { e->reason = R_FAULT; e->fault_addr = f.addr; return R_FAULT; }
}
// resume after a Python hook handled the call at hook_pc: the handler sets x0 etc.; this returns to the link register
void emu_return_from_hook(Emu* e) { e->pc = e->x[30]; }
// execute the instruction at pc even though it carries a Python hook (used by "observe and continue" hooks)
int emu_step_over(Emu* e) { try { if (!step(e)) return e->reason; e->icount++; return -1; } catch (Fault& f) { e->reason = R_FAULT; e->fault_addr = f.addr; return R_FAULT; } }

}  // extern "C"
