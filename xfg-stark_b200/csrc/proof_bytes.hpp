// proof_bytes.hpp — host-side serialisation helpers shared by the tuned pipeline (prover.cu), the general-options pipeline
// (general_api.inc) and its CPU emulation harness (tests/host_emul): byte writer, `BatchMerkleProof::serialize_nodes`, coin seed elements,
// and `StarkProof::to_bytes` for the general-options state.  Plain C++ (no CUDA).
//
// Replaces winter-crypto 0.8.3 `MerkleTree::prove_batch` + `BatchMerkleProof::serialize_nodes` and winter-air `StarkProof::to_bytes`
//  as reached from `air.prove(trace)` (src/burn_mint_prover.rs:124) -> `proof.to_bytes()` (src/bin/xfg-stark-cli.rs:533).
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>
#include "../../include/xfg_stark.h"
#include "field.cuh"

namespace xfg {

// coin seed elements: Context::to_elements() then the public inputs (A.4)
inline void seed_elements(u32 ln, const xfg_options& o, u32 width, const u64* pub_inputs, u32 num_pub, u64* out) {
  int k = 0;
  out[k++] = (u64)width << 8;
  out[k++] = XFG_P & 0xFFFFFFFFull; out[k++] = XFG_P >> 32;
  out[k++] = (u64)o.field_extension << 16 | (u64)o.fri_folding_factor << 8 | o.fri_remainder_max_degree;
  out[k++] = o.grinding_factor; out[k++] = o.blowup_factor; out[k++] = o.num_queries;
  out[k++] = (u64)(u32)(size_t(1) << ln);
  for (u32 i = 0; i < num_pub; i++) out[k++] = pub_inputs[i];
}

struct Out { std::vector<u8> b;
  void u8_(u32 v) { b.push_back((u8)v); } void u16_(size_t v) { b.push_back((u8)v); b.push_back((u8)(v >> 8)); }
  void u32_(size_t v) { for (int i = 0; i < 4; i++) b.push_back((u8)(v >> (8 * i))); } void u64_(u64 v) { for (int i = 0; i < 8; i++) b.push_back((u8)(v >> (8 * i))); }
  void raw(const void* p, size_t n) { const u8* q = (const u8*)p; b.insert(b.end(), q, q + n); } };

// BatchMerkleProof::serialize_nodes of MerkleTree::prove_batch(positions) (A.11), built from the per-position sibling paths:
// path(q, lvl) = tree[((M + pos[q]) >> lvl) ^ 1].  Follows the crate's bookkeeping exactly (norm = sorted unique (index & ~1); at every level
// `nodes[i]` is indexed by the position i in the current node list), without its BTreeMap: the ancestors (M + pos) >> lvl =
// (M >> lvl) + (pos >> lvl) are monotone in pos, so the owner of a node is found by binary search over the positions sorted once
// (any queried leaf below a node yields the same sibling digest).  ~10 us for 42 positions in a 2^23-leaf tree (the linear-search version: 35 us,
// seven trees per proof).
inline void batch_paths(const u32* pos, u32 cnt, const u64* paths, u32 depth, u64 M, Out& out) {
  auto path = [&](u32 q, u32 lvl) { return reinterpret_cast<const u8*>(paths + ((size_t)q * depth + lvl) * 4); };
  u32 ord[256]; for (u32 q = 0; q < cnt; q++) ord[q] = q;
  std::sort(ord, ord + cnt, [&](u32 a, u32 b) { return pos[a] < pos[b]; });
  auto owner = [&](u64 heap_index, u32 lvl) -> int {
    const u64 key = heap_index - (M >> lvl); u32 lo = 0, hi = cnt;
    while (lo < hi) { const u32 mid = (lo + hi) / 2; if (((u64)pos[ord[mid]] >> lvl) < key) lo = mid + 1; else hi = mid; }
    return (lo < cnt && ((u64)pos[ord[lo]] >> lvl) == key) ? (int)ord[lo] : -1;
  };
  u64 cur[256], next[256]; u32 nn = 0, nnext = 0;
  static thread_local std::vector<const u8*> flat; flat.resize((size_t)256 * (depth + 2)); u32 ncount[256];
  const size_t stride = depth + 2;
  u64 norm[256];
  for (u32 i = 0; i < cnt; i++) { const u64 v = pos[ord[i]] & ~u64(1); if (!nn || norm[nn - 1] != v) norm[nn++] = v; }
  for (u32 k = 0; k < nn; k++) {
    const u64 index = norm[k]; ncount[k] = 0;
    for (u64 i = index; i < index + 2; i++) if (owner(M + i, 0) < 0) flat[k * stride + ncount[k]++] = path((u32)owner(M + (i ^ 1), 0), 0);
    next[nnext++] = (index + M) >> 1;
  }
  for (u32 d = 1; d < depth; d++) {
    const u32 nc = nnext; for (u32 i = 0; i < nc; i++) cur[i] = next[i];
    nnext = 0;
    u32 i = 0;
    while (i < nc) {
      const u64 sib = cur[i] ^ 1;
      if (i + 1 < nc && cur[i + 1] == sib) i += 1;
      else flat[i * stride + ncount[i]++] = path((u32)owner(cur[i], d), d);
      next[nnext++] = sib >> 1; i += 1;
    }
  }
  size_t total = 1; for (u32 k = 0; k < nn; k++) total += 1 + 32 * (size_t)ncount[k];
  const size_t base = out.b.size(); out.b.resize(base + total);
  u8* w = out.b.data() + base; *w++ = (u8)nn;
  for (u32 k = 0; k < nn; k++) { *w++ = (u8)ncount[k]; for (u32 j = 0; j < ncount[k]; j++) { std::memcpy(w, flat[k * stride + j], 32); w += 32; } }
}
}  // namespace xfg
