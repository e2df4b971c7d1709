// CTC prefix beam search on the GPU (no language model) — semantics per SURVEY.md Appendix D / oracle/beam.py.
// PARITY UNPINNED: the reference delegates to the external paddlespeech_ctcdecoders C++ library
// (masr/decoders/swig_wrapper.py:35-64, beam_search_decoder.py:45-56), which is absent here.
//
//   ctc_topk_kernel        per frame: the `cutoff_top_n` most probable tokens (descending, ties -> lower id), truncated
//                          where the cumulative probability reaches `cutoff_prob`; emits ids + log-probabilities.
//                          The [T,V] posterior never goes to the host (the reference ships it as Python lists).
//   prefix_beam_kernel     one CTA per utterance walks the frames; the beam lives in shared memory, the prefix trie
//                          (parent, token) plus a persistent (parent, token) -> node hash in global memory, so a prefix that
//                          drops out of the beam and is re-created later keeps its identity (and its children in the beam
//                          keep merging with it) exactly like the restatement's `child` dictionary — without it two beam
//                          entries could spell the same prefix and split its mass (seen as a ln 2 score gap on a 12 s
//                          utterance); selection = exact radix select + bitonic sort, so ties
//                          resolve deterministically (existing prefixes by rank, then children in (parent rank, candidate)
//                          order) and the result equals the CPU restatement.
#include <math.h>

#include "common.cuh"

namespace masr {

constexpr int BK_MAX = 40;        // cutoff_top_n cap
constexpr int BEAM_CAP = 512;     // beam_size cap (reference default 300)
constexpr int BEAM_THREADS = 512;

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ctc_topk_kernel(const float* __restrict__ logits, int64_t ldl, int V, int top_n,
                                                       float cutoff_prob, int* __restrict__ cand_id,
                                                       float* __restrict__ cand_logp, int* __restrict__ cand_cnt) {
    constexpr int PER = 20;                    // 256 * 20 >= 4233 (larger vocabularies take the strided fallback below)
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* x = logits + (int64_t)row * ldl;
    __shared__ float s_val[8];
    __shared__ int s_idx[8];
    __shared__ float s_red[8];
    __shared__ float s_pick_v[BK_MAX];
    __shared__ int s_pick_i[BK_MAX];
    float v[PER];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int i = tid + j * 256;
        v[j] = i < V ? __ldg(x + i) : -INFINITY;
        mx = fmaxf(mx, v[j]);
    }
    for (int i = tid + PER * 256; i < V; i += 256) mx = fmaxf(mx, __ldg(x + i));   // only if V > 5120
    mx = warp_max(mx);
    if ((tid & 31) == 0) s_red[tid >> 5] = mx;
    __syncthreads();
    mx = s_red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, s_red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int i = tid; i < V; i += 256) sum += expf(__ldg(x + i) - mx);
    sum = warp_sum(sum);
    if ((tid & 31) == 0) s_red[tid >> 5] = sum;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_red[w];
    // top_n rounds of block arg-max over the register-resident values
    for (int r = 0; r < top_n; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = tid + j * 256;
            if (v[j] > bv) { bv = v[j]; bi = i; }      // ascending i within a thread: strict > keeps the lowest index
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { s_val[tid >> 5] = bv; s_idx[tid >> 5] = bi; }
        __syncthreads();
        bv = s_val[0]; bi = s_idx[0];
#pragma unroll
        for (int w = 1; w < 8; ++w)
            if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bi)) { bv = s_val[w]; bi = s_idx[w]; }
        if (tid == 0) { s_pick_v[r] = bv; s_pick_i[r] = bi; }
        if ((bi & 255) == tid) {                                               // owner retires the winner
#pragma unroll
            for (int j = 0; j < PER; ++j)
                if (j == (bi >> 8)) v[j] = -INFINITY;
        }
        __syncthreads();
    }
    if (tid == 0) {
        float cum = 0.f;
        int n = 0;
        for (int r = 0; r < top_n; ++r) {
            if (s_pick_v[r] == -INFINITY) break;
            const float p = expf(s_pick_v[r] - mx) / tot;       // the float32 posterior the reference would pass
            cand_id[(int64_t)row * BK_MAX + n] = s_pick_i[r];
            cand_logp[(int64_t)row * BK_MAX + n] = logf(p);
            ++n;
            cum += p;
            if (cum >= cutoff_prob) break;
        }
        cand_cnt[row] = n;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// log(exp(a) + exp(b)) in a SPECIFIED sequence of correctly rounded float32 operations (no FMA contraction, no libm): the
// same sequence as oracle/beam.py's exp32_det / log1p32_det, so kernel and restatement agree bit for bit — a pruned search over
// hundreds of frames turns any 1-ulp score difference into a different beam.
__device__ __forceinline__ float exp_det(float d) {                    // d <= 0
    if (d < -87.0f) return 0.f;
    const float n = rintf(__fmul_rn(d, 1.4426950408889634f));
    float r = __fsub_rn(d, __fmul_rn(n, 0.693145751953125f));
    r = __fsub_rn(r, __fmul_rn(n, 1.42860682030941723212e-6f));
    float p = 1.0f / 720.0f;
    p = __fadd_rn(__fmul_rn(p, r), 1.0f / 120.0f);
    p = __fadd_rn(__fmul_rn(p, r), 1.0f / 24.0f);
    p = __fadd_rn(__fmul_rn(p, r), 1.0f / 6.0f);
    p = __fadd_rn(__fmul_rn(p, r), 0.5f);
    p = __fadd_rn(__fmul_rn(p, r), 1.0f);
    p = __fadd_rn(__fmul_rn(p, r), 1.0f);
    return __fmul_rn(p, __int_as_float(((int)n + 127) << 23));          // * 2^n (n in [-126, 0])
}
__device__ __forceinline__ float log1p_det(float u) {                  // 0 <= u <= 1: 2 atanh(u / (2 + u))
    const float s = __fdiv_rn(u, __fadd_rn(2.0f, u));
    const float z = __fmul_rn(s, s);
    float p = 2.0f / 17.0f;
    p = __fadd_rn(__fmul_rn(p, z), 2.0f / 15.0f);
    p = __fadd_rn(__fmul_rn(p, z), 2.0f / 13.0f);
    p = __fadd_rn(__fmul_rn(p, z), 2.0f / 11.0f);
    p = __fadd_rn(__fmul_rn(p, z), 2.0f / 9.0f);
    p = __fadd_rn(__fmul_rn(p, z), 2.0f / 7.0f);
    p = __fadd_rn(__fmul_rn(p, z), 2.0f / 5.0f);
    p = __fadd_rn(__fmul_rn(p, z), 2.0f / 3.0f);
    p = __fadd_rn(__fmul_rn(p, z), 2.0f);
    return __fmul_rn(s, p);
}
__device__ __forceinline__ float logaddexp_f(float a, float b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    const float hi = fmaxf(a, b), lo = fminf(a, b);
    return __fadd_rn(hi, log1p_det(exp_det(__fsub_rn(lo, hi))));
}
// order-preserving float -> uint key (larger float -> larger key); -inf maps lowest
__device__ __forceinline__ uint32_t fkey(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct BeamShared {
    int node[BEAM_CAP], par[BEAM_CAP], last[BEAM_CAP];
    float pb[BEAM_CAP], pnb[BEAM_CAP], score[BEAM_CAP];
    float nb[BEAM_CAP], nnb[BEAM_CAP];          // next-frame accumulators of the existing prefixes
    int hkey[2 * BEAM_CAP], hval[2 * BEAM_CAP]; // node id -> beam index
    int s_node[BEAM_CAP], s_par[BEAM_CAP], s_last[BEAM_CAP];   // staging for the re-ranked beam
    float s_pb[BEAM_CAP], s_pnb[BEAM_CAP], s_score[BEAM_CAP];
    int s_src[BEAM_CAP];                        // pool index of each survivor
    uint32_t hist[256];
    int cid[BK_MAX];
    float clp[BK_MAX];
    int scan[BEAM_THREADS];
    int misc[8];
    int wsum[2][BEAM_THREADS / 32];              // per-warp (gt | eq << 16) counts of the ordered compaction, double-buffered
    float r_score[BEAM_CAP];                     // survivors in rank order
    int r_src[BEAM_CAP];
};

// pool layout: [0, BEAM_CAP) existing prefixes (rank order), then BEAM_CAP + i*K + k children of (rank i, candidate k);
// K = this frame's candidate count (usually a handful, cutoff_prob 0.99), so the pool the selection scans is 512 + beam*K
// entries, not 512 + beam*40
__global__ void __launch_bounds__(BEAM_THREADS) prefix_beam_kernel(
    const int* __restrict__ cand_id, const float* __restrict__ cand_logp, const int* __restrict__ cand_cnt, int64_t bstride,
    const int* __restrict__ lens, int beam, int blank, float* __restrict__ pool_all, int* __restrict__ trie_parent,
    int* __restrict__ trie_tok, int64_t trie_cap, int* __restrict__ out_tok, int64_t tok_stride, int* __restrict__ out_n,
    float* __restrict__ out_score, int* __restrict__ state_i, float* __restrict__ state_f, int resume) {
    // state_i / state_f (optional, per utterance 3*BEAM_CAP+2 ints / 3*BEAM_CAP floats): the beam after the last frame, so the
    // search can be resumed with the next chunk of frames (`resume` != 0) — CTCBeamSearchDecoder.next()/decode() of the
    // reference's streaming path (beam_search_decoder.py:75-91); the trie and its hash persist in trie_parent / trie_tok.
    extern __shared__ __align__(16) uint8_t smem_beam[];
    BeamShared& S = *reinterpret_cast<BeamShared*>(smem_beam);
    const int b = blockIdx.x, tid = threadIdx.x;
    const int T = lens[b];
    float* pool = pool_all + (int64_t)b * (BEAM_CAP + BEAM_CAP * BK_MAX);
    int* tpar = trie_parent + (int64_t)b * trie_cap;
    int* ttok = trie_tok + (int64_t)b * trie_cap;
    // per-utterance trie storage: [0, node_cap) nodes, then in `trie_parent` an open-addressing hash of node ids keyed by
    // (parent, token) (4 slots per possible node; compared through tpar/ttok of the stored id)
    const int64_t node_cap = trie_cap / 5;
    const uint32_t hcap = (uint32_t)(trie_cap - node_cap);
    int* thash = tpar + node_cap;
    int nbeam = 1, nnodes = 1;
    int* st_i = state_i ? state_i + (int64_t)b * (3 * BEAM_CAP + 2) : nullptr;
    float* st_f = state_f ? state_f + (int64_t)b * (3 * BEAM_CAP) : nullptr;
    if (resume && st_i) {
        nbeam = st_i[3 * BEAM_CAP];
        nnodes = st_i[3 * BEAM_CAP + 1];
        if (tid < nbeam) {
            S.node[tid] = st_i[tid]; S.par[tid] = st_i[BEAM_CAP + tid]; S.last[tid] = st_i[2 * BEAM_CAP + tid];
            S.pb[tid] = st_f[tid]; S.pnb[tid] = st_f[BEAM_CAP + tid]; S.score[tid] = st_f[2 * BEAM_CAP + tid];
        }
    } else {
        for (int64_t i = tid; i < hcap; i += BEAM_THREADS) thash[i] = -1;
        if (tid == 0) {
            S.node[0] = 0; S.par[0] = -1; S.last[0] = -1; S.pb[0] = 0.f; S.pnb[0] = -INFINITY; S.score[0] = 0.f;
            tpar[0] = -1; ttok[0] = -1;
        }
    }
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        const int64_t row = (int64_t)b * bstride + t;
        const int K = cand_cnt[row];
        if (tid < K) { S.cid[tid] = cand_id[row * BK_MAX + tid]; S.clp[tid] = cand_logp[row * BK_MAX + tid]; }
        for (int i = tid; i < 2 * BEAM_CAP; i += BEAM_THREADS) S.hkey[i] = -1;
        const int pool_n = BEAM_CAP + nbeam * K;
        for (int i = tid; i < pool_n; i += BEAM_THREADS) pool[i] = -INFINITY;
        __syncthreads();
        // node id -> rank hash; stay transitions (blank / repeated token) of the existing prefixes
        if (tid < nbeam) {
            uint32_t h = ((uint32_t)S.node[tid] * 2654435761u) & (2 * BEAM_CAP - 1);
            while (atomicCAS(&S.hkey[h], -1, S.node[tid]) != -1) h = (h + 1) & (2 * BEAM_CAP - 1);
            S.hval[h] = tid;
            float nb = -INFINITY, nnb = -INFINITY;
            for (int k = 0; k < K; ++k) {
                if (S.cid[k] == blank) nb = logaddexp_f(nb, S.score[tid] + S.clp[k]);
                else if (S.cid[k] == S.last[tid]) nnb = logaddexp_f(nnb, S.pnb[tid] + S.clp[k]);
            }
            S.nb[tid] = nb; S.nnb[tid] = nnb;
        }
        __syncthreads();
        // extensions: child (rank i, candidate k)
        for (int e = tid; e < nbeam * K; e += BEAM_THREADS) {
            const int i = e / K, k = e - i * K;
            const int c = S.cid[k];
            if (c == blank) continue;
            float add;
            if (c == S.last[i]) add = S.pb[i] == -INFINITY ? -INFINITY : S.pb[i] + S.clp[k];
            else add = S.score[i] + S.clp[k];
            pool[BEAM_CAP + i * K + k] = add;
        }
        __syncthreads();
        // children that already exist as beam entries: fold their contribution into that entry (one pair per entry)
        if (tid < nbeam && S.par[tid] >= 0) {
            uint32_t h = ((uint32_t)S.par[tid] * 2654435761u) & (2 * BEAM_CAP - 1);
            int pi = -1;
            while (S.hkey[h] != -1) {
                if (S.hkey[h] == S.par[tid]) { pi = S.hval[h]; break; }
                h = (h + 1) & (2 * BEAM_CAP - 1);
            }
            if (pi >= 0) {
                for (int k = 0; k < K; ++k)
                    if (S.cid[k] == S.last[tid]) {
                        const int slot = BEAM_CAP + pi * K + k;
                        S.nnb[tid] = logaddexp_f(S.nnb[tid], pool[slot]);
                        pool[slot] = -INFINITY;
                        break;
                    }
            }
        }
        __syncthreads();
        if (tid < nbeam) pool[tid] = logaddexp_f(S.nb[tid], S.nnb[tid]);
        __syncthreads();
        // ---- exact top-`beam` selection over the pool: 4-pass radix select on the order-preserving key ----
        uint32_t prefix = 0, mask = 0;
        int want = beam;
        for (int pass = 3; pass >= 0; --pass) {
            for (int i = tid; i < 256; i += BEAM_THREADS) S.hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < pool_n; i += BEAM_THREADS) {
                const uint32_t kx = fkey(pool[i]);
                if ((kx & mask) == prefix && pool[i] != -INFINITY) atomicAdd(&S.hist[(kx >> (pass * 8)) & 255], 1u);
            }
            __syncthreads();
            // the digit where the descending cumulative count reaches `want`: inclusive scan over the 256 bins by 8 warps
            // (the first version walked the bins serially in thread 0 — a third of the kernel's time in the r02 capture)
            {
                int v = 0, incl = 0;
                if (tid < 256) {
                    v = (int)S.hist[255 - tid];                     // position tid <-> digit 255 - tid
                    incl = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int u = __shfl_up_sync(0xffffffffu, incl, o);
                        if ((tid & 31) >= o) incl += u;
                    }
                    if ((tid & 31) == 31) S.wsum[0][tid >> 5] = incl;
                }
                __syncthreads();
                if (tid < 256) {
                    for (int w = 0; w < (tid >> 5); ++w) incl += S.wsum[0][w];
                    const int excl = incl - v;
                    if (excl < want && (incl >= want || tid == 255)) { S.misc[0] = 255 - tid; S.misc[1] = want - excl; }
                }
            }
            __syncthreads();
            prefix |= (uint32_t)S.misc[0] << (pass * 8);
            mask |= 0xFFu << (pass * 8);
            want = S.misc[1];
            __syncthreads();
        }
        // threshold key = prefix; take everything above it, and the first `want` (pool order) equal to it.
        // Ordered compaction: per chunk of 512 pool entries one ballot per warp + the 16 warp counts through shared memory
        // (double-buffered: one block barrier per chunk; the first version ran a 9-step shared-memory scan per chunk).
        const int lane_ = tid & 31, warp_ = tid >> 5;
        const unsigned lt_mask = (1u << lane_) - 1u;
        int run_gt = 0, run_eq = 0;                                             // identical in every thread
        int chunk = 0;
        for (int start = 0; start < pool_n; start += BEAM_THREADS, ++chunk) {
            const int i = start + tid;
            int is_gt = 0, is_eq = 0;
            if (i < pool_n && pool[i] != -INFINITY) {
                const uint32_t kx = fkey(pool[i]);
                is_gt = kx > prefix;
                is_eq = kx == prefix;
            }
            const unsigned bg = __ballot_sync(0xffffffffu, is_gt), be = __ballot_sync(0xffffffffu, is_eq);
            if (lane_ == 0) S.wsum[chunk & 1][warp_] = __popc(bg) | (__popc(be) << 16);
            __syncthreads();
            int off_g = 0, off_e = 0, tot_g = 0, tot_e = 0;
#pragma unroll
            for (int w = 0; w < BEAM_THREADS / 32; ++w) {
                const int v = S.wsum[chunk & 1][w];
                if (w < warp_) { off_g += v & 0xFFFF; off_e += v >> 16; }
                tot_g += v & 0xFFFF; tot_e += v >> 16;
            }
            const int gt_before = run_gt + off_g + __popc(bg & lt_mask);
            const int eq_before = run_eq + off_e + __popc(be & lt_mask);
            if (is_gt) S.s_src[gt_before] = i;                                  // provisional: gt entries first
            if (is_eq && eq_before < want) S.s_src[BEAM_CAP - 1 - eq_before] = i; // eq entries parked at the tail
            run_gt += tot_g; run_eq += tot_e;
        }
        __syncthreads();
        const int n_gt = run_gt;
        const int n_eq = min(run_eq, want);
        const int n_sel = n_gt + n_eq;
        if (tid < n_eq) S.s_score[tid] = __int_as_float(S.s_src[BEAM_CAP - 1 - tid]);   // (staged: the tail may overlap [n_gt, n_sel))
        __syncthreads();
        if (tid < n_eq) S.s_src[n_gt + tid] = __float_as_int(S.s_score[tid]);
        __syncthreads();
        // ---- rank the survivors by (score desc, pool index asc): bitonic sort of the 512 (score, index) pairs held one per
        //      thread; compare-exchange distances below 32 go through warp shuffles, only the 10 distances >= 32 through shared
        //      memory (the first version: 45 block barriers; rank-by-counting was tried in between and was slower) ----
        {
            float ks = -INFINITY;
            int ki = 0x7fffffff;
            if (tid < n_sel) { ki = S.s_src[tid]; ks = pool[ki]; }
            for (int size = 2; size <= BEAM_CAP; size <<= 1) {
                const bool up = (tid & size) == 0;
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    float c;
                    int ci;
                    if (stride >= 32) {
                        __syncthreads();                         // (previous readers of the exchange buffers are done)
                        S.s_score[tid] = ks; S.scan[tid] = ki;
                        __syncthreads();
                        c = S.s_score[tid ^ stride]; ci = S.scan[tid ^ stride];
                    } else {
                        c = __shfl_xor_sync(0xffffffffu, ks, stride);
                        ci = __shfl_xor_sync(0xffffffffu, ki, stride);
                    }
                    const bool mine_first = (ks > c) || (ks == c && ki < ci);
                    const bool want_first = ((tid & stride) == 0) == up;
                    if (mine_first != want_first) { ks = c; ki = ci; }
                }
            }
            __syncthreads();
            if (tid < n_sel) { S.r_score[tid] = ks; S.r_src[tid] = ki; }
            __syncthreads();
        }
        // ---- materialise the new beam ----
        if (tid < n_sel) {
            const int src = S.r_src[tid];
            if (src < BEAM_CAP) {                     // an existing prefix survives
                S.s_node[tid] = S.node[src]; S.s_par[tid] = S.par[src]; S.s_last[tid] = S.last[src];
                S.s_pb[tid] = S.nb[src]; S.s_pnb[tid] = S.nnb[src];
                S.s_src[tid] = -1;
            } else {                                  // a new child: gets a trie node below
                const int i = (src - BEAM_CAP) / K, k = (src - BEAM_CAP) - i * K;
                S.s_node[tid] = -1; S.s_par[tid] = S.node[i]; S.s_last[tid] = S.cid[k];
                S.s_pb[tid] = -INFINITY; S.s_pnb[tid] = pool[src];
                S.s_src[tid] = 1;
            }
        }
        __syncthreads();
        // new children: reuse the node of a prefix that existed before (persistent hash), else allocate ids in rank order
        int need_new = 0;
        if (tid < n_sel && S.s_src[tid] == 1) {
            const int par = S.s_par[tid], tok = S.s_last[tid];
            uint32_t h = (((uint32_t)par * 2654435761u) ^ ((uint32_t)tok * 40503u)) % hcap;
            int found = -1;
            for (;;) {
                const int id = thash[h];
                if (id == -1) break;
                if (tpar[id] == par && ttok[id] == tok) { found = id; break; }
                h = h + 1 == hcap ? 0 : h + 1;
            }
            S.s_node[tid] = found;
            need_new = found < 0;
        }
        int new_total;
        {
            const unsigned bn = __ballot_sync(0xffffffffu, need_new);
            if (lane_ == 0) S.wsum[0][warp_] = __popc(bn);
            __syncthreads();
            int off = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < BEAM_THREADS / 32; ++w) { const int v = S.wsum[0][w]; if (w < warp_) off += v; tot += v; }
            S.scan[tid] = off + __popc(bn & lt_mask) + need_new;              // inclusive count, as before
            new_total = tot;
        }
        if (need_new) {
            const int id = nnodes + S.scan[tid] - 1;          // rank order (deterministic)
            S.s_node[tid] = id;
            if (id < node_cap) {
                const int par = S.s_par[tid], tok = S.s_last[tid];
                tpar[id] = par; ttok[id] = tok;
                uint32_t h = (((uint32_t)par * 2654435761u) ^ ((uint32_t)tok * 40503u)) % hcap;
                while (atomicCAS(&thash[h], -1, id) != -1) h = h + 1 == hcap ? 0 : h + 1;
            }
        }
        __syncthreads();
        nnodes += new_total;
        if (tid < n_sel) {
            S.node[tid] = S.s_node[tid]; S.par[tid] = S.s_par[tid]; S.last[tid] = S.s_last[tid];
            S.pb[tid] = S.s_pb[tid]; S.pnb[tid] = S.s_pnb[tid]; S.score[tid] = S.r_score[tid];
        }
        nbeam = n_sel;
        __syncthreads();
        if (nbeam == 0) break;
    }
    if (st_i) {
        if (tid < nbeam) {
            st_i[tid] = S.node[tid]; st_i[BEAM_CAP + tid] = S.par[tid]; st_i[2 * BEAM_CAP + tid] = S.last[tid];
            st_f[tid] = S.pb[tid]; st_f[BEAM_CAP + tid] = S.pnb[tid]; st_f[2 * BEAM_CAP + tid] = S.score[tid];
        }
        if (tid == 0) { st_i[3 * BEAM_CAP] = nbeam; st_i[3 * BEAM_CAP + 1] = nnodes; }
    }
    if (tid == 0) {
        int n = 0;
        float sc = -INFINITY;
        if (nbeam > 0) {
            sc = S.score[0];
            int node = S.node[0];
            int len = 0;
            for (int x = node; x > 0 && x < node_cap; x = tpar[x]) ++len;
            n = len;
            int pos = len - 1;
            for (int x = node; x > 0 && x < node_cap && pos >= 0; x = tpar[x]) out_tok[(int64_t)b * tok_stride + pos--] = ttok[x];
        }
        out_n[b] = n;
        out_score[b] = sc;
    }
}

}  // namespace masr

using namespace masr;

extern "C" int masr_ctc_topk_f32(const float* logits, int64_t ldl, int M, int V, int top_n, float cutoff_prob, int* cand_id,
                                 float* cand_logp, int* cand_cnt, void* stream) {
    if (M == 0) return MASR_OK;
    MASR_REQUIRE(logits && cand_id && cand_logp && cand_cnt, "masr_ctc_topk_f32: null pointer");
    MASR_REQUIRE(top_n >= 1 && top_n <= BK_MAX, "masr_ctc_topk_f32: cutoff_top_n=%d out of range (1..%d)", top_n, BK_MAX);
    MASR_REQUIRE(V <= 20 * 256, "masr_ctc_topk_f32: vocabulary %d > 5120 not supported by this build", V);
    ctc_topk_kernel<<<M, 256, 0, (cudaStream_t)stream>>>(logits, ldl, V, top_n, cutoff_prob, cand_id, cand_logp, cand_cnt);
    return check_launch("ctc_topk_kernel");
}

extern "C" int masr_ctc_prefix_beam_workspace(int B, int Tmax, int64_t* pool_floats, int64_t* trie_ints_per_utt) {
    MASR_REQUIRE(pool_floats && trie_ints_per_utt, "masr_ctc_prefix_beam_workspace: null pointer");
    *pool_floats = (int64_t)B * (BEAM_CAP + BEAM_CAP * BK_MAX);
    *trie_ints_per_utt = 5 * ((int64_t)Tmax * BEAM_CAP + 1);      // nodes + 4 hash slots per possible node
    return MASR_OK;
}

extern "C" int masr_ctc_prefix_beam(const int* cand_id, const float* cand_logp, const int* cand_cnt, int64_t bstride,
                                    const int* lens, int B, int beam_size, int blank, float* pool, int* trie_parent,
                                    int* trie_tok, int64_t trie_cap, int* out_tok, int64_t tok_stride, int* out_n,
                                    float* out_score, void* stream) {
    if (B == 0) return MASR_OK;
    MASR_REQUIRE(cand_id && cand_logp && cand_cnt && lens && pool && trie_parent && trie_tok && out_tok && out_n && out_score,
                 "masr_ctc_prefix_beam: null pointer");
    MASR_REQUIRE(beam_size >= 1 && beam_size <= BEAM_CAP, "masr_ctc_prefix_beam: beam_size=%d out of range (1..%d)", beam_size, BEAM_CAP);
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(prefix_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BeamShared));
        if (e != cudaSuccess) { set_last_error("prefix_beam smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_set[dev] = true;
    }
    prefix_beam_kernel<<<B, BEAM_THREADS, sizeof(BeamShared), (cudaStream_t)stream>>>(
        cand_id, cand_logp, cand_cnt, bstride, lens, beam_size, blank, pool, trie_parent, trie_tok, trie_cap, out_tok, tok_stride,
        out_n, out_score, nullptr, nullptr, 0);
    return check_launch("prefix_beam_kernel");
}

extern "C" int masr_ctc_prefix_beam_state_size(int64_t* ints_per_utt, int64_t* floats_per_utt) {
    MASR_REQUIRE(ints_per_utt && floats_per_utt, "masr_ctc_prefix_beam_state_size: null pointer");
    *ints_per_utt = 3 * BEAM_CAP + 2;
    *floats_per_utt = 3 * BEAM_CAP;
    return MASR_OK;
}

// Streaming form (beam_search_decoder.py:75-96: CTCBeamSearchDecoder.next() + decode(), reset_state()): the same search
// fed chunk by chunk.  `lens[b]` = frames of THIS chunk (0 = no new frames for that stream), `resume` = 0 starts a new
// utterance (reset_decoder), != 0 continues from `state_*`; trie_parent / trie_tok must be sized for the whole stream
// (masr_ctc_prefix_beam_workspace with Tmax = the longest stream in frames) and persist between calls.  Outputs = the best
// prefix and its score after the frames seen so far — identical to one masr_ctc_prefix_beam call over the concatenation.
extern "C" int masr_ctc_prefix_beam_stream(const int* cand_id, const float* cand_logp, const int* cand_cnt, int64_t bstride,
                                           const int* lens, int B, int beam_size, int blank, float* pool, int* trie_parent,
                                           int* trie_tok, int64_t trie_cap, int* state_i, float* state_f, int resume,
                                           int* out_tok, int64_t tok_stride, int* out_n, float* out_score, void* stream) {
    if (B == 0) return MASR_OK;
    MASR_REQUIRE(cand_id && cand_logp && cand_cnt && lens && pool && trie_parent && trie_tok && out_tok && out_n && out_score &&
                 state_i && state_f, "masr_ctc_prefix_beam_stream: null pointer");
    MASR_REQUIRE(beam_size >= 1 && beam_size <= BEAM_CAP, "masr_ctc_prefix_beam_stream: beam_size=%d out of range (1..%d)", beam_size, BEAM_CAP);
    cudaError_t e = cudaFuncSetAttribute(prefix_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BeamShared));
    if (e != cudaSuccess) { set_last_error("prefix_beam smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    prefix_beam_kernel<<<B, BEAM_THREADS, sizeof(BeamShared), (cudaStream_t)stream>>>(
        cand_id, cand_logp, cand_cnt, bstride, lens, beam_size, blank, pool, trie_parent, trie_tok, trie_cap, out_tok, tok_stride,
        out_n, out_score, state_i, state_f, resume);
    return check_launch("prefix_beam_kernel<stream>");
}
