// reverb_b200 — CTC head post-processing and searches on the GPU.
//
//   logsoftmax_topk : per frame log_softmax over V (reference: transformer/ctc.py:106-114) fused with the top-N
//                     selection that both searches start from (transformer/search.py:111,155) — the (B,T',V) log-prob
//                     tensor (1.9 GB at B=64) is only written when the caller asks for it.
//   ctc_greedy      : arg-max path, padded frames -> blank, collapse repeats, drop blanks (search.py:106-121,
//                     utils/ctc_utils.py:22-32).
//   ctc_prefix_beam : CTC prefix beam search with the reference's exact update rules, float64 score arithmetic and
//                     Viterbi time tracking (search.py:124-248, utils/common.py:355-363) — one warp per utterance,
//                     prefixes as canonical trie nodes, times as persistent linked lists.
//   logsoftmax_gather : rescoring decoder output -> log-probs of the hypothesis tokens only (search.py:413-436).
#include <math.h>

#include "kernels.h"

namespace rvb {

// ---------------------------------------------------------------------------------------------------------------
struct ArgMax {
  float v;
  int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ ArgMax warp_argmax(ArgMax a) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ArgMax b;
    b.v = __shfl_xor_sync(0xffffffffu, a.v, o);
    b.i = __shfl_xor_sync(0xffffffffu, a.i, o);
    a = better(a, b);
  }
  return a;
}

__global__ void __launch_bounds__(256)
logsoftmax_topk_kernel(const float* __restrict__ logits, long long ld, int V, int k, float* __restrict__ topk_val,
                       int* __restrict__ topk_idx, float* __restrict__ logp_out, int apply_softmax) {
  extern __shared__ float s_row[];  // V floats
  __shared__ float s_red[8];
  __shared__ ArgMax s_arg[8];
  __shared__ float s_stat[2];
  const long long row = blockIdx.x;
  const float* x = logits + row * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += 256) {
    float v = x[i];
    s_row[i] = v;
    m = fmaxf(m, v);
  }
  float lse_shift = 0.f, logsum = 0.f;
  if (apply_softmax) {
    m = warp_max(m);
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    m = s_red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < V; i += 256) s += expf(s_row[i] - m);
    s = warp_sum(s);
    if (lane == 0) s_red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += s_red[w];
      s_stat[0] = m;
      s_stat[1] = logf(t);
    }
    __syncthreads();
    lse_shift = s_stat[0];
    logsum = s_stat[1];
  } else {
    __syncthreads();
  }
  if (logp_out != nullptr) {
    for (int i = threadIdx.x; i < V; i += 256) logp_out[row * V + i] = (s_row[i] - lse_shift) - logsum;
  }
  // top-k by repeated block arg-max (ties -> lowest index); values reported as log-probs
  for (int r = 0; r < k; ++r) {
    ArgMax a;
    a.v = -INFINITY;
    a.i = 0x7fffffff;
    for (int i = threadIdx.x; i < V; i += 256) {
      ArgMax b;
      b.v = s_row[i];
      b.i = i;
      a = better(a, b);
    }
    a = warp_argmax(a);
    if (lane == 0) s_arg[warp] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
      ArgMax best = s_arg[0];
      for (int w = 1; w < 8; ++w) best = better(best, s_arg[w]);
      if (best.i == 0x7fffffff) best.i = 0;  // fewer than k finite entries
      topk_val[row * k + r] = (best.v - lse_shift) - logsum;
      topk_idx[row * k + r] = best.i;
      s_row[best.i] = -INFINITY;
      // NaN-safe marker: a -inf entry can be re-selected only when everything left is -inf
    }
    __syncthreads();
  }
}

int launch_logsoftmax_topk(const float* logits, int ld, int M, int V, int k, float* topk_val, int* topk_idx,
                           float* logp_out, int apply_softmax, cudaStream_t stream) {
  RVB_REQUIRE(k >= 1 && k <= 16 && k <= V, "logsoftmax_topk: k=%d unsupported", k);
  if (M <= 0) return 0;
  const size_t smem = (size_t)V * sizeof(float);
  RVB_REQUIRE(smem <= 200 * 1024, "logsoftmax_topk: V=%d too large for the shared-memory row cache", V);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    RVB_CHECK_CUDA(cudaFuncSetAttribute(logsoftmax_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  logsoftmax_topk_kernel<<<M, 256, smem, stream>>>(logits, ld, V, k, topk_val, topk_idx, logp_out, apply_softmax);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void ctc_greedy_kernel(const int* __restrict__ top1, int stride, const int* __restrict__ lens, int T,
                                  int blank, int* __restrict__ out_tokens, int* __restrict__ out_lens) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int len = min(lens[b], T);
  int prev = -1, count = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    int t = t0 + lane;
    int id = blank;
    if (t < len) id = top1[((long long)b * T + t) * stride];
    int left = __shfl_up_sync(0xffffffffu, id, 1);
    if (lane == 0) left = prev;
    bool keep = (t < T) && (id != blank) && (id != left);
    unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) out_tokens[(long long)b * T + count + __popc(mask & ((1u << lane) - 1u))] = id;
    count += __popc(mask);
    prev = __shfl_sync(0xffffffffu, id, 31);
  }
  if (lane == 0) out_lens[b] = count;
}

int launch_ctc_greedy(const int* top1_idx, int idx_stride, const int* lens, int B, int T, int blank, int* out_tokens,
                      int* out_lens, cudaStream_t stream) {
  if (B <= 0) return 0;
  ctc_greedy_kernel<<<B, 32, 0, stream>>>(top1_idx, idx_stride, lens, T, blank, out_tokens, out_lens);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// CTC prefix beam search.
constexpr int PB_MAXBEAM = 16;
constexpr int PB_MAXSLOTS = PB_MAXBEAM + PB_MAXBEAM * PB_MAXBEAM;
#define PB_NEG_INF (-INFINITY)

__device__ __forceinline__ double log_add2(double a, double b) {
  // utils/common.py:355-363 for two arguments
  if (a == PB_NEG_INF) return b;
  if (b == PB_NEG_INF) return a;
  double mx = a > b ? a : b;
  double sum = exp(a - mx) + exp(b - mx);
  return mx + log(sum);
}

struct PBLayout {
  size_t per_utt_ints;
  int pool_cap, hash_cap, times_cap;
};
__host__ __device__ inline PBLayout pb_layout(int T, int beam) {
  PBLayout l;
  l.pool_cap = beam * T + 2;
  int h = 64;
  while (h < 2 * l.pool_cap) h <<= 1;
  l.hash_cap = h;
  l.times_cap = 2 * beam * T + 2;
  l.per_utt_ints = (size_t)2 * l.pool_cap + l.hash_cap + (size_t)2 * l.times_cap;
  return l;
}
size_t prefix_beam_workspace_bytes(int B, int T, int beam) {
  PBLayout l = pb_layout(T, beam);
  return l.per_utt_ints * sizeof(int) * (size_t)B;
}

struct PBSlot {
  double s, ns, vs, vns, ctp;
  int times_s;           // times list node (-1 = empty list)
  int tns_src, tns_op;   // times_ns recipe: op 0 = none, 1 = append t to list tns_src, 2 = replace last of tns_src by t
  int src, tok;          // extension slots: source beam entry and token; stay slots: src = own index, tok = -1
};

constexpr int PB_THREADS = 128;
constexpr int PB_KEY_NONE = 0x7fffffff;

// One CTA (4 warps) per utterance.  Per frame:
//   P0  stage the frame's top-k, derive score()/viterbi_score()/times() of every beam prefix
//   P1  "stay" slots (prefix unchanged): at most three updates land on prefix j — blank, repeat of its last token,
//       and the one extension parent(j)+last(j) that equals j — replayed in the reference's iteration order;
//       extension slots: one thread per (token, prefix) pair, final at once unless it collides with a stay slot
//   P2  score() of every touched slot          P3  rank (score desc, dict insertion order asc), keep top `beam`,
//       materialise survivors: canonical trie node (find-or-create in a shared-memory hash), times list nodes
// Prefix identity = canonical trie node id, so dict-merge semantics of the reference hold exactly.
__global__ void __launch_bounds__(PB_THREADS)
ctc_prefix_beam_kernel(const float* __restrict__ topk_val, const int* __restrict__ topk_idx, int k,
                       const int* __restrict__ lens, int T, int beam, int blank, int* __restrict__ workspace,
                       int trie_in_smem, int max_len, int* __restrict__ out_tokens, int* __restrict__ out_times,
                       int* __restrict__ out_lens, double* __restrict__ out_scores, int* __restrict__ out_nhyp) {
  extern __shared__ int pb_dyn[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const PBLayout L = pb_layout(T, beam);
  int* ws = workspace + (size_t)b * L.per_utt_ints;
  int* trie_parent = trie_in_smem ? pb_dyn : ws;
  int* trie_tok = trie_parent + L.pool_cap;
  int* hash = trie_parent + 2 * L.pool_cap;  // -1 = empty
  int* times_parent = ws + 2 * L.pool_cap + L.hash_cap;
  int* times_t = times_parent + L.times_cap;

  __shared__ PBSlot slots[PB_MAXSLOTS];
  __shared__ double slot_score[PB_MAXSLOTS];
  __shared__ int slot_key[PB_MAXSLOTS];
  __shared__ double c_s[PB_MAXBEAM], c_ns[PB_MAXBEAM], c_vs[PB_MAXBEAM], c_vns[PB_MAXBEAM];
  __shared__ double c_score[PB_MAXBEAM], c_vit[PB_MAXBEAM];
  __shared__ int c_node[PB_MAXBEAM], c_ts[PB_MAXBEAM], c_tns[PB_MAXBEAM], c_tnsp[PB_MAXBEAM], c_times[PB_MAXBEAM];
  __shared__ int c_last[PB_MAXBEAM], c_par[PB_MAXBEAM];
  __shared__ double n_s[PB_MAXBEAM], n_ns[PB_MAXBEAM], n_vs[PB_MAXBEAM], n_vns[PB_MAXBEAM];
  __shared__ double n_score[PB_MAXBEAM];
  __shared__ int n_node[PB_MAXBEAM], n_ts[PB_MAXBEAM], n_tns[PB_MAXBEAM], n_tnsp[PB_MAXBEAM];
  __shared__ double s_tv[PB_MAXBEAM];
  __shared__ int s_ti[PB_MAXBEAM];
  __shared__ int s_pool, s_times;

  if (trie_in_smem)
    for (int i = tid; i < L.hash_cap; i += PB_THREADS) hash[i] = -1;
  if (tid == 0) {
    trie_parent[0] = -1;
    trie_tok[0] = -1;
    s_pool = 1;
    s_times = 0;
    n_node[0] = 0;
    n_s[0] = 0.0;
    n_ns[0] = PB_NEG_INF;
    n_vs[0] = 0.0;
    n_vns[0] = 0.0;
    n_score[0] = 0.0;  // log_add([0, -inf])
    n_ts[0] = -1;
    n_tns[0] = -1;
    n_tnsp[0] = -1;
  }
  int nb = 1;
  const int len = min(lens[b], T);
  const int kk = min(k, beam);
  __syncthreads();

  for (int t = 0; t < len; ++t) {
    // ---- P0
    if (tid < kk) {
      s_tv[tid] = (double)topk_val[((long long)b * T + t) * k + tid];
      s_ti[tid] = topk_idx[((long long)b * T + t) * k + tid];
    }
    if (tid < nb) {
      const int j = tid;
      const double s = n_s[j], ns = n_ns[j], vs = n_vs[j], vns = n_vns[j];
      c_s[j] = s;
      c_ns[j] = ns;
      c_vs[j] = vs;
      c_vns[j] = vns;
      c_ts[j] = n_ts[j];
      c_tns[j] = n_tns[j];
      c_tnsp[j] = n_tnsp[j];
      const int node = n_node[j];
      c_node[j] = node;
      c_score[j] = n_score[j];  // == log_add([s, ns]); already computed when slot j was ranked
      const bool sb = vs > vns;
      c_vit[j] = sb ? vs : vns;
      c_times[j] = sb ? n_ts[j] : n_tns[j];
      c_last[j] = trie_tok[node];
      c_par[j] = trie_parent[node];
      PBSlot& sl = slots[j];
      sl.s = sl.ns = sl.vs = sl.vns = sl.ctp = PB_NEG_INF;
      sl.times_s = -1;
      sl.tns_src = -1;
      sl.tns_op = 0;
      sl.src = j;
      sl.tok = -1;
      slot_key[j] = PB_KEY_NONE;
    }
    __syncthreads();
    const int ncand = kk * nb;
    const int nslots = nb + ncand;
    // ---- P1a: stay slots
    if (tid < nb) {
      const int j = tid;
      PBSlot& n = slots[j];
      const int last_j = c_last[j];
      const bool nonempty = c_node[j] != 0;
      int ui_blank = -1, ui_last = -1, ip = -1;
      for (int u = 0; u < kk; ++u) {
        if (s_ti[u] == blank) ui_blank = u;
        if (nonempty && s_ti[u] == last_j) ui_last = u;
      }
      if (nonempty && ui_last >= 0)
        for (int i = 0; i < nb; ++i)
          if (c_node[i] == c_par[j]) ip = i;
      // events sorted by candidate index c = ui * nb + i
      int ev_c[3], ev_type[3], ne = 0;  // type 0 blank, 1 repeat, 2 collision
      if (ui_blank >= 0) { ev_c[ne] = ui_blank * nb + j; ev_type[ne++] = 0; }
      if (ui_last >= 0) { ev_c[ne] = ui_last * nb + j; ev_type[ne++] = 1; }
      if (ui_last >= 0 && ip >= 0) { ev_c[ne] = ui_last * nb + ip; ev_type[ne++] = 2; }
      for (int a = 1; a < ne; ++a)
        for (int q = a; q > 0 && ev_c[q] < ev_c[q - 1]; --q) {
          int tc = ev_c[q]; ev_c[q] = ev_c[q - 1]; ev_c[q - 1] = tc;
          int tt = ev_type[q]; ev_type[q] = ev_type[q - 1]; ev_type[q - 1] = tt;
        }
      int key = PB_KEY_NONE;
      for (int a = 0; a < ne; ++a) {
        const int ty = ev_type[a];
        if (ty == 0) {
          const double p = s_tv[ui_blank];
          if (key == PB_KEY_NONE) key = 2 * ev_c[a];
          n.s = log_add2(n.s, c_score[j] + p);
          n.vs = c_vit[j] + p;
          n.times_s = c_times[j];
        } else if (ty == 1) {
          const double p = s_tv[ui_last];
          if (key == PB_KEY_NONE) key = 2 * ev_c[a];
          n.ns = log_add2(n.ns, c_ns[j] + p);
          if (n.vns < c_vns[j] + p) {
            // reference typo (`vs_ns`, search.py:178): v_ns is NOT updated here
            if (n.ctp < p) {
              n.ctp = p;
              n.tns_src = c_tnsp[j];  // parent of prefix j's times_ns list: "copy, then overwrite the last element"
              n.tns_op = 2;
            }
          }
        } else {
          const double p = s_tv[ui_last];
          if (key == PB_KEY_NONE) key = 2 * ev_c[a] + 1;
          const bool rep = (last_j == c_last[ip]) && (c_node[ip] != 0);
          const double add = rep ? c_s[ip] : c_score[ip];
          const double vit = rep ? c_vs[ip] : c_vit[ip];
          n.ns = log_add2(n.ns, add + p);
          if (n.vns < vit + p) {
            n.vns = vit + p;
            n.ctp = p;
            n.tns_src = rep ? c_ts[ip] : c_times[ip];
            n.tns_op = 1;
          }
        }
      }
      slot_key[j] = key;
    }
    // ---- P1b: extension slots, one thread per (token, prefix) candidate
    for (int c = tid; c < ncand; c += PB_THREADS) {
      const int ui = c / nb, i = c - ui * nb;
      const int u = s_ti[ui];
      const double p = s_tv[ui];
      int key = PB_KEY_NONE;
      if (u != blank) {
        const int node_i = c_node[i];
        bool collide = false;
        for (int j = 0; j < nb; ++j) collide |= (c_par[j] == node_i && c_last[j] == u && c_node[j] != 0);
        if (!collide) {
          PBSlot& e = slots[nb + c];
          const bool rep = (u == c_last[i]) && (node_i != 0);
          // a single contribution into a fresh PrefixScore (s = ns = v_s = v_ns = -inf): log_add([-inf, x]) == x
          const double add = rep ? c_s[i] : c_score[i];
          const double vit = rep ? c_vs[i] : c_vit[i];
          e.s = PB_NEG_INF;
          e.ns = add + p;
          e.vs = PB_NEG_INF;
          e.vns = PB_NEG_INF;
          e.ctp = PB_NEG_INF;
          e.times_s = -1;
          e.tns_src = -1;
          e.tns_op = 0;
          if (PB_NEG_INF < vit + p) {
            e.vns = vit + p;
            e.ctp = p;
            e.tns_src = rep ? c_ts[i] : c_times[i];
            e.tns_op = 1;
          }
          e.src = i;
          e.tok = u;
          key = 2 * c + 1;
        }
      }
      slot_key[nb + c] = key;
    }
    __syncthreads();
    // ---- P2: score() of every touched slot (extension slots have s = -inf: no transcendental)
    for (int a = tid; a < nslots; a += PB_THREADS)
      slot_score[a] = (slot_key[a] == PB_KEY_NONE) ? PB_NEG_INF : log_add2(slots[a].s, slots[a].ns);
    __syncthreads();
    // ---- P3: second beam prune (stable w.r.t. dict insertion order) + materialise the survivors
    int nlive = 0;
    for (int o = 0; o < nslots; ++o) nlive += (slot_key[o] != PB_KEY_NONE);
    const int nnew = min(beam, nlive);
    for (int a = tid; a < nslots; a += PB_THREADS) {
      const int key = slot_key[a];
      if (key == PB_KEY_NONE) continue;
      const double sc = slot_score[a];
      int rank = 0;
      for (int o = 0; o < nslots; ++o) {
        const int ko = slot_key[o];
        const double so = slot_score[o];
        rank += (ko != PB_KEY_NONE) && (so > sc || (so == sc && ko < key));
      }
      if (rank >= nnew) continue;
      const PBSlot& s = slots[a];
      int node;
      if (s.tok < 0) {
        node = c_node[s.src];
      } else {
        const int parent = c_node[s.src];
        unsigned h = ((unsigned)parent * 2654435761u) ^ ((unsigned)s.tok * 40503u + 0x9e3779b9u);
        h &= (unsigned)(L.hash_cap - 1);
        node = -1;
        int fresh = -1;
        while (true) {
          int cur = atomicAdd(&hash[h], 0);
          if (cur == -1) {
            if (fresh < 0) {
              fresh = atomicAdd(&s_pool, 1);
              trie_parent[fresh] = parent;
              trie_tok[fresh] = s.tok;
              __threadfence_block();
            }
            int old = atomicCAS(&hash[h], -1, fresh);
            if (old == -1) {
              node = fresh;
              break;
            }
            cur = old;
          }
          if (trie_parent[cur] == parent && trie_tok[cur] == s.tok) {
            node = cur;
            break;
          }
          h = (h + 1) & (unsigned)(L.hash_cap - 1);
        }
      }
      int tns = -1, tnsp = -1;
      if (s.tns_op != 0) {
        tns = atomicAdd(&s_times, 1);
        tnsp = s.tns_src;      // op 1: append to list tns_src; op 2: tns_src already is the parent to hang t on
        times_parent[tns] = tnsp;
        times_t[tns] = t;
      }
      n_node[rank] = node;
      n_s[rank] = s.s;
      n_ns[rank] = s.ns;
      n_vs[rank] = s.vs;
      n_vns[rank] = s.vns;
      n_score[rank] = sc;
      n_ts[rank] = s.times_s;
      n_tns[rank] = tns;
      n_tnsp[rank] = tnsp;
    }
    nb = nnew;
    __syncthreads();
  }

  // ---- emit the n-best: tokens, score() and times() per surviving prefix, in beam order
  if (tid == 0) out_nhyp[b] = nb;
  if (tid < nb) {
    const int r = tid;
    int n = 0;
    for (int node = n_node[r]; node > 0; node = trie_parent[node]) ++n;
    int* tok_out = out_tokens + ((long long)b * beam + r) * max_len;
    int* tim_out = out_times + ((long long)b * beam + r) * max_len;
    int pos = n;
    for (int node = n_node[r]; node > 0; node = trie_parent[node]) {
      --pos;
      if (pos < max_len) tok_out[pos] = trie_tok[node];
    }
    const int tl = (n_vs[r] > n_vns[r]) ? n_ts[r] : n_tns[r];
    int nt = 0;
    for (int q = tl; q >= 0; q = times_parent[q]) ++nt;
    pos = nt;
    for (int q = tl; q >= 0; q = times_parent[q]) {
      --pos;
      if (pos < max_len) tim_out[pos] = times_t[q];
    }
    out_lens[(b * beam + r) * 2 + 0] = n;
    out_lens[(b * beam + r) * 2 + 1] = nt;
    out_scores[b * beam + r] = n_score[r];
  }
}

int launch_ctc_prefix_beam(const float* topk_val, const int* topk_idx, int k, const int* lens, int B, int T, int beam,
                           int blank, void* workspace, size_t workspace_bytes, int max_len, int* out_tokens,
                           int* out_times, int* out_lens, double* out_scores, int* out_nhyp, cudaStream_t stream) {
  RVB_REQUIRE(beam >= 1 && beam <= PB_MAXBEAM, "prefix beam: beam_size=%d unsupported (1..%d)", beam, PB_MAXBEAM);
  RVB_REQUIRE(k >= beam, "prefix beam: need top-k with k >= beam (k=%d beam=%d)", k, beam);
  const size_t need = prefix_beam_workspace_bytes(B, T, beam);
  RVB_REQUIRE(workspace_bytes >= need, "prefix beam: workspace too small (%zu < %zu)", workspace_bytes, need);
  if (B <= 0) return 0;
  const PBLayout L = pb_layout(T, beam);
  const size_t trie_bytes = ((size_t)2 * L.pool_cap + L.hash_cap) * sizeof(int);
  const int trie_in_smem = trie_bytes <= 150 * 1024;
  if (!trie_in_smem) RVB_CHECK_CUDA(cudaMemsetAsync(workspace, 0xFF, need, stream));  // hash tables = -1
  const size_t dyn = trie_in_smem ? trie_bytes : 0;
  static size_t configured = 0;
  if (dyn > configured) {
    RVB_CHECK_CUDA(cudaFuncSetAttribute(ctc_prefix_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    configured = dyn;
  }
  ctc_prefix_beam_kernel<<<B, PB_THREADS, dyn, stream>>>(topk_val, topk_idx, k, lens, T, beam, blank,
                                                        reinterpret_cast<int*>(workspace), trie_in_smem, max_len,
                                                        out_tokens, out_times, out_lens, out_scores, out_nhyp);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
logsoftmax_gather_kernel(const float* __restrict__ logits, long long ld, int V, const int* __restrict__ gidx, int G,
                         float* __restrict__ out) {
  __shared__ float s_red[8];
  __shared__ float s_stat[2];
  const long long row = blockIdx.x;
  const float* x = logits + row * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < V; i += 256) m = fmaxf(m, x[i]);
  m = warp_max(m);
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  m = s_red[0];
  for (int w = 1; w < 8; ++w) m = fmaxf(m, s_red[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += 256) s += expf(x[i] - m);
  s = warp_sum(s);
  if (lane == 0) s_red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    s_stat[0] = m;
    s_stat[1] = logf(t);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < G; j += 256) {
    int id = gidx[row * G + j];
    out[row * G + j] = (id >= 0 && id < V) ? (x[id] - s_stat[0]) - s_stat[1] : 0.f;
  }
}

int launch_logsoftmax_gather(const float* logits, int ld, int M, int V, const int* gather_idx, int G, float* out,
                             cudaStream_t stream) {
  if (M <= 0) return 0;
  logsoftmax_gather_kernel<<<M, 256, 0, stream>>>(logits, ld, V, gather_idx, G, out);
  RVB_COUNT_LAUNCH();
  RVB_CHECK_LAUNCH();
  return 0;
}

}  // namespace rvb
