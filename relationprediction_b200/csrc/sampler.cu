// sampler.cu -- host-side neighbourhood-expansion edge sampler (next row N2).
//
// Same stochastic process as the reference's sample_edge_neighborhood (train.py:161-198):
//   repeat sample_size times:
//     pick a vertex v with probability proportional to (#unpicked incident edges of v) * seen[v]
//       (if that mass is zero: uniformly among vertices that still have unpicked incident edges);
//     pick one of v's unpicked incident edges uniformly (the reference does it by rejection sampling over
//       the adjacency list, which is the same distribution);
//     mark the edge picked, both endpoints seen, decrement both endpoints' counts.
// The reference draws the vertex with np.random.choice over a V-long probability vector every iteration
// (O(V) per draw, ~5 s for 30 000 edges on FB15k-237); here the weights live in two Fenwick trees and the
// unpicked incident edges in swap-remove lists, O(log V) per draw (~10 ms for the same sample).
// The random stream differs from numpy's, so parity is distributional (tests compare statistics).
#include <stdint.h>

#include <algorithm>
#include <new>
#include <random>
#include <string>
#include <vector>

#include "graph.h"

namespace {

struct Fenwick {
  std::vector<int64_t> t;
  int n;
  explicit Fenwick(int n_) : t((size_t)n_ + 1, 0), n(n_) {}
  void add(int i, int64_t v) {
    for (++i; i <= n; i += i & -i) t[i] += v;
  }
  int64_t total() const {
    int64_t s = 0;
    for (int i = n; i > 0; i -= i & -i) s += t[i];
    return s;
  }
  // smallest index i with prefix_sum(i) > r   (0 <= r < total)
  int find(int64_t r) const {
    int pos = 0, lg = 1;
    while ((lg << 1) <= n) lg <<= 1;
    for (int k = lg; k > 0; k >>= 1) {
      if (pos + k <= n && t[pos + k] <= r) {
        pos += k;
        r -= t[pos];
      }
    }
    return pos;
  }
};

}  // namespace

// Immutable per-dataset part: incidence lists in their pristine order.  A draw copies the mutable arrays
// (~20 B per edge, a memcpy) instead of rebuilding them, which was more than half of a 30 000-edge sample.
struct rgcn_sampler {
  int64_t E = 0;
  int32_t V = 0;
  std::vector<int32_t> sub, obj;             // endpoints of every edge
  std::vector<int64_t> off;                  // V+1 list offsets
  std::vector<int32_t> inc_edge, inc_other;  // 2E entries: every edge once per endpoint
  std::vector<int32_t> pos_s, pos_o;         // where the edge sits in its subject's / object's list
  std::vector<int64_t> any_tree;             // Fenwick tree over [degree > 0]
};

namespace {

int sampler_build(const int32_t* triples_host, int64_t E, int32_t V, rgcn_sampler& sp) {
  if (!triples_host || E <= 0 || V <= 0 || 2 * E >= (int64_t)1 << 31) {
    rgcn_set_error("edge sampler: need E > 0, V > 0 and fewer than 2^30 edges");
    return RGCN_ERR_INVALID;
  }
  sp.E = E;
  sp.V = V;
  sp.sub.resize((size_t)E);
  sp.obj.resize((size_t)E);
  sp.off.assign((size_t)V + 1, 0);
  for (int64_t e = 0; e < E; ++e) {
    const int32_t s = triples_host[3 * e], o = triples_host[3 * e + 2];
    if (s < 0 || s >= V || o < 0 || o >= V) {
      rgcn_set_error("edge sampler: entity id out of range");
      return RGCN_ERR_INVALID;
    }
    sp.sub[e] = s;
    sp.obj[e] = o;
    sp.off[s + 1]++;
    sp.off[o + 1]++;
  }
  for (int32_t v = 0; v < V; ++v) sp.off[v + 1] += sp.off[v];
  sp.inc_edge.resize((size_t)2 * E);
  sp.inc_other.resize((size_t)2 * E);
  sp.pos_s.resize((size_t)E);
  sp.pos_o.resize((size_t)E);
  std::vector<int64_t> cur(sp.off.begin(), sp.off.end() - 1);
  for (int64_t e = 0; e < E; ++e) {
    const int32_t s = sp.sub[e], o = sp.obj[e];
    sp.pos_s[e] = (int32_t)cur[s]++;
    sp.inc_edge[sp.pos_s[e]] = (int32_t)e;
    sp.inc_other[sp.pos_s[e]] = o;
    sp.pos_o[e] = (int32_t)cur[o]++;
    sp.inc_edge[sp.pos_o[e]] = (int32_t)e;
    sp.inc_other[sp.pos_o[e]] = s;
  }
  Fenwick any(V);
  for (int32_t v = 0; v < V; ++v)
    if (sp.off[v + 1] > sp.off[v]) any.add(v, 1);
  sp.any_tree = any.t;
  return RGCN_OK;
}

int sampler_draw(const rgcn_sampler& sp, int64_t sample_size, uint64_t seed, int32_t* out_edges_host) {
  const int64_t E = sp.E;
  const int32_t V = sp.V;
  if (!out_edges_host || sample_size < 0 || sample_size > E) {
    rgcn_set_error("edge sampler: need 0 <= sample_size <= E");
    return RGCN_ERR_INVALID;
  }
  const std::vector<int64_t>& off = sp.off;
  std::vector<int32_t> inc_edge(sp.inc_edge), inc_other(sp.inc_other), pos_s(sp.pos_s), pos_o(sp.pos_o);
  std::vector<int64_t> live((size_t)V);  // number of unpicked entries at the front of each list
  for (int32_t v = 0; v < V; ++v) live[v] = off[v + 1] - off[v];
  std::vector<char> seen((size_t)V, 0);
  Fenwick w_seen(V);  // weight = live[v] if seen[v] else 0
  Fenwick w_any(V);   // weight = 1 if live[v] > 0
  w_any.t = sp.any_tree;

  std::mt19937_64 rng(seed);
  auto uniform = [&](int64_t n) { return (int64_t)(rng() % (uint64_t)n); };

  // remove the list entry at absolute position p of vertex v (swap with the last live entry)
  auto remove_entry = [&](int32_t v, int64_t p) {
    const int64_t last = off[v] + live[v] - 1;
    if (p != last) {
      const int32_t e2 = inc_edge[last];
      std::swap(inc_edge[p], inc_edge[last]);
      std::swap(inc_other[p], inc_other[last]);
      // the moved entry belongs to edge e2: fix whichever of its two positions pointed at `last`
      if (pos_s[e2] == last && sp.sub[e2] == v)
        pos_s[e2] = (int32_t)p;
      else
        pos_o[e2] = (int32_t)p;
    }
    live[v]--;
    if (seen[v]) w_seen.add(v, -1);
    if (live[v] == 0) w_any.add(v, -1);
  };
  auto mark_seen = [&](int32_t v) {
    if (!seen[v]) {
      seen[v] = 1;
      if (live[v] > 0) w_seen.add(v, live[v]);
    }
  };

  for (int64_t i = 0; i < sample_size; ++i) {
    int32_t v;
    const int64_t tot = w_seen.total();
    if (tot > 0) {
      v = w_seen.find(uniform(tot));
    } else {
      const int64_t alive = w_any.total();
      if (alive <= 0) {
        rgcn_set_error("edge sampler: ran out of edges");
        return RGCN_ERR_INVALID;
      }
      v = w_any.find(uniform(alive));
    }
    mark_seen(v);
    const int64_t p = off[v] + uniform(live[v]);
    const int32_t e = inc_edge[p];
    const int32_t other = inc_other[p];
    out_edges_host[i] = e;
    // drop the edge from both endpoints' lists (a self loop sits twice in the same list)
    const int32_t s = sp.sub[e], o = sp.obj[e];
    if (s == o) {
      // two entries in v's list: remove the one at the larger position first so the other index stays valid
      const int64_t a = std::max(pos_s[e], pos_o[e]), b = std::min(pos_s[e], pos_o[e]);
      remove_entry(s, a);
      remove_entry(s, b);
    } else {
      remove_entry(s, pos_s[e]);
      remove_entry(o, pos_o[e]);
    }
    mark_seen(other);
  }
  return RGCN_OK;
}

}  // namespace

extern "C" int rgcn_sampler_create(const int32_t* triples_host, int64_t E, int32_t V, rgcn_sampler** out) {
  if (!out) {
    rgcn_set_error("rgcn_sampler_create: null output");
    return RGCN_ERR_INVALID;
  }
  *out = nullptr;
  rgcn_sampler* sp = new (std::nothrow) rgcn_sampler();
  if (!sp) return RGCN_ERR_NOMEM;
  try {
    const int rc = sampler_build(triples_host, E, V, *sp);
    if (rc != RGCN_OK) {
      delete sp;
      return rc;
    }
  } catch (const std::bad_alloc&) {
    delete sp;
    rgcn_set_error("rgcn_sampler_create: out of host memory");
    return RGCN_ERR_NOMEM;
  }
  *out = sp;
  return RGCN_OK;
}

// thread-compatible: concurrent draws on one handle are fine (the handle is read-only, each draw owns its copies)
extern "C" int rgcn_sampler_draw(const rgcn_sampler* sp, int64_t sample_size, uint64_t seed,
                                 int32_t* out_edges_host) {
  if (!sp) {
    rgcn_set_error("rgcn_sampler_draw: null handle");
    return RGCN_ERR_INVALID;
  }
  try {
    return sampler_draw(*sp, sample_size, seed, out_edges_host);
  } catch (const std::bad_alloc&) {
    rgcn_set_error("rgcn_sampler_draw: out of host memory");
    return RGCN_ERR_NOMEM;
  }
}

extern "C" void rgcn_sampler_destroy(rgcn_sampler* sp) { delete sp; }

extern "C" int rgcn_sample_edge_neighborhood(const int32_t* triples_host, int64_t E, int32_t V,
                                             int64_t sample_size, uint64_t seed,
                                             int32_t* out_edges_host) {
  if (!triples_host || !out_edges_host || E <= 0 || V <= 0 || sample_size < 0 || sample_size > E) {
    rgcn_set_error("rgcn_sample_edge_neighborhood: need 0 <= sample_size <= E, E > 0, V > 0");
    return RGCN_ERR_INVALID;
  }
  rgcn_sampler sp;
  try {
    const int rc = sampler_build(triples_host, E, V, sp);
    if (rc != RGCN_OK) return rc;
    return sampler_draw(sp, sample_size, seed, out_edges_host);
  } catch (const std::bad_alloc&) {
    rgcn_set_error("rgcn_sample_edge_neighborhood: out of host memory");
    return RGCN_ERR_NOMEM;
  }
}
