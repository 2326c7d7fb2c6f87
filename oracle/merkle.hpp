// oracle/merkle.hpp — TEST INFRASTRUCTURE (CPU oracle). Not part of the product.
//
// Restates winter-crypto 0.8.3 `MerkleTree::{new, prove_batch}`, `BatchMerkleProof::{serialize_nodes, deserialize,
// get_root}` (reached from DefaultTraceLde::new, src/burn_mint_air.rs:513).
#pragma once
#include <map>
#include <set>
#include <stdexcept>
#include "hash.hpp"

namespace orc {

extern int g_threads;

struct BatchMerkleProof {
  std::vector<Digest> leaves;               // in the order of the queried indexes
  std::vector<std::vector<Digest>> nodes;   // one vector per normalised (even) index
  unsigned depth = 0;

  // u8 count, then per vector u8 len + digests (A.11)
  std::vector<u8> serialize_nodes() const {
    std::vector<u8> r;
    if (nodes.size() > 255) throw std::runtime_error("too many paths");
    r.push_back((u8)nodes.size());
    for (auto& v : nodes) {
      if (v.size() > 255) throw std::runtime_error("too many nodes");
      r.push_back((u8)v.size());
      for (auto& d : v) r.insert(r.end(), d.begin(), d.end());
    }
    return r;
  }
  static bool deserialize(const u8* p, size_t len, std::vector<Digest> leaves, unsigned depth, BatchMerkleProof& out) {
    size_t pos = 0; if (len < 1) return false;
    size_t nv = p[pos++]; out.nodes.clear();
    for (size_t i = 0; i < nv; i++) {
      if (pos >= len) return false;
      size_t k = p[pos++]; if (pos + 32 * k > len) return false;
      std::vector<Digest> v(k);
      for (size_t j = 0; j < k; j++) { std::memcpy(v[j].data(), p + pos, 32); pos += 32; }
      out.nodes.push_back(std::move(v));
    }
    if (pos != len) return false;
    out.leaves = std::move(leaves); out.depth = depth; return true;
  }
  // winter-crypto BatchMerkleProof::get_root; returns false on a malformed proof
  bool get_root(const std::vector<size_t>& indexes_in, Digest& root) const;
};

inline bool map_indexes(const std::vector<size_t>& indexes, unsigned depth, std::map<size_t, size_t>& m) {
  size_t num_leaves = size_t(1) << depth;
  for (size_t i = 0; i < indexes.size(); i++) {
    if (indexes[i] >= num_leaves) return false;
    if (!m.emplace(indexes[i], i).second) return false;  // duplicate
  }
  return true;
}
inline std::vector<size_t> normalize_indexes(const std::vector<size_t>& indexes) {
  std::set<size_t> s; for (size_t i : indexes) s.insert(i - (i & 1));
  return std::vector<size_t>(s.begin(), s.end());
}

struct MerkleTree {
  std::vector<Digest> leaves, nodes;  // nodes[1] = root, nodes[0] unused
  explicit MerkleTree(std::vector<Digest> lv) : leaves(std::move(lv)) {
    size_t n = leaves.size();
    if (n < 2 || (n & (n - 1))) throw std::runtime_error("number of leaves must be a power of two >= 2");
    nodes.assign(n, Digest{});
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && n >= 4096)
    for (size_t i = 0; i < n / 2; i++) nodes[n / 2 + i] = merge(leaves[2 * i], leaves[2 * i + 1]);
    for (size_t lo = n / 4; lo >= 1; lo >>= 1) {   // level [lo, 2*lo): same values as the serial i = n/2-1 .. 1 loop
#pragma omp parallel for num_threads(g_threads) schedule(static) if (g_threads > 1 && lo >= 2048)
      for (size_t i = lo; i < 2 * lo; i++) nodes[i] = merge(nodes[2 * i], nodes[2 * i + 1]);
    }
  }
  const Digest& root() const { return nodes[1]; }
  unsigned depth() const { return ilog2_(leaves.size()); }
  static unsigned ilog2_(size_t n) { unsigned k = 0; while ((size_t(1) << k) < n) k++; return k; }

  BatchMerkleProof prove_batch(const std::vector<size_t>& indexes_in) const {
    if (indexes_in.empty()) throw std::runtime_error("TooFewLeafIndexes");
    if (indexes_in.size() > 255) throw std::runtime_error("TooManyLeafIndexes");
    std::map<size_t, size_t> index_map;
    if (!map_indexes(indexes_in, depth(), index_map)) throw std::runtime_error("bad leaf indexes");
    std::vector<size_t> indexes = normalize_indexes(indexes_in);
    BatchMerkleProof pf; pf.depth = depth();
    pf.leaves.assign(index_map.size(), Digest{});
    size_t n = leaves.size();
    std::vector<size_t> next;
    for (size_t index : indexes) {
      std::vector<Digest> missing;
      for (size_t i = index; i < index + 2; i++) {
        auto it = index_map.find(i);
        if (it != index_map.end()) pf.leaves[it->second] = leaves[i]; else missing.push_back(leaves[i]);
      }
      pf.nodes.push_back(std::move(missing));
      next.push_back((index + n) >> 1);
    }
    for (unsigned d = 1; d < depth(); d++) {
      std::vector<size_t> cur = next; next.clear();
      size_t i = 0;
      while (i < cur.size()) {
        size_t sib = cur[i] ^ 1;
        if (i + 1 < cur.size() && cur[i + 1] == sib) i += 1;
        else pf.nodes[i].push_back(nodes[sib]);   // indexed by position i in `cur` (A.11, as in the reference crate)
        next.push_back(sib >> 1);
        i += 1;
      }
    }
    return pf;
  }
};

inline bool BatchMerkleProof::get_root(const std::vector<size_t>& indexes_in, Digest& root) const {
  if (indexes_in.empty() || indexes_in.size() > 255) return false;
  std::map<size_t, size_t> index_map;
  if (!map_indexes(indexes_in, depth, index_map)) return false;
  std::vector<size_t> indexes = normalize_indexes(indexes_in);
  if (indexes.size() != nodes.size()) return false;
  std::map<size_t, Digest> v;
  size_t offset = size_t(1) << depth;
  std::vector<size_t> next, ptr;
  for (size_t i = 0; i < indexes.size(); i++) {
    size_t index = indexes[i];
    Digest b0, b1;
    auto i1 = index_map.find(index), i2 = index_map.find(index + 1);
    if (i1 != index_map.end()) {
      if (leaves.size() <= i1->second) return false;
      b0 = leaves[i1->second];
      if (i2 != index_map.end()) { if (leaves.size() <= i2->second) return false; b1 = leaves[i2->second]; ptr.push_back(0); }
      else { if (nodes[i].empty()) return false; b1 = nodes[i][0]; ptr.push_back(1); }
    } else {
      if (nodes[i].empty()) return false;
      b0 = nodes[i][0];
      if (i2 == index_map.end()) return false;
      if (leaves.size() <= i2->second) return false;
      b1 = leaves[i2->second]; ptr.push_back(1);
    }
    size_t parent = (offset + index) >> 1;
    v[parent] = merge(b0, b1); next.push_back(parent);
  }
  for (unsigned d = 1; d < depth; d++) {
    std::vector<size_t> cur = next; next.clear();
    size_t i = 0;
    while (i < cur.size()) {
      size_t node_index = cur[i], sib_index = node_index ^ 1;
      Digest sib;
      if (i + 1 < cur.size() && cur[i + 1] == sib_index) {
        auto it = v.find(sib_index); if (it == v.end()) return false; sib = it->second; i += 1;
      } else {
        size_t p = ptr[i]; if (nodes[i].size() <= p) return false;
        sib = nodes[i][p]; ptr[i] += 1;
      }
      auto itn = v.find(node_index); if (itn == v.end()) return false;
      Digest parent = (node_index & 1) ? merge(sib, itn->second) : merge(itn->second, sib);
      v[node_index >> 1] = parent; next.push_back(node_index >> 1);
      i += 1;
    }
  }
  auto it = v.find(1); if (it == v.end()) return false;
  root = it->second; return true;
}

}  // namespace orc
