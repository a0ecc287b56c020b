// On-device beam-search bookkeeping of Seq2seq._step_translate (reference models/Seq2seq.py:337-393): two kernels replace the
// ~25 ATen launches per decode position (log_softmax, topk, masked_fill x2, div, topk, mul, floor-div, remainder, gathers,
// index copies, eos / length updates, sum) of the step that b200st.decode.BeamSearch replays.
//   topk_logsoftmax: one pass over a row of logits -> log-sum-exp and the k best entries (k <= 8), scores = logit - lse (fp32).
//   beam_select:     per utterance, the k*k candidates (beam r, choice j) are scored with the reference's rule
//                    (hypothesis score + log-probability, finished beams keep only their first candidate at +0, division
//                    by len^alpha), the k best are taken, and the token prefixes / KV-cache ancestry / key masks are
//                    re-ordered in the same launch; EOS flags, lengths and the all-finished counter follow.
#include "common.cuh"

namespace b200st {

constexpr int BK_MAX = 8;

template <typename T>
__global__ void __launch_bounds__(256) topk_logsoftmax_kernel(const T* __restrict__ x, int64_t ld, int cols, int k,
                                                              float* __restrict__ score, int64_t* __restrict__ pred) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float sv[8][BK_MAX];
  __shared__ int si[8][BK_MAX];
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const T* xr = x + r * ld;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // thread-local top-k (descending; ties keep the lower index), running max / sum-exp
  float tv[BK_MAX];
  int ti[BK_MAX];
#pragma unroll
  for (int q = 0; q < BK_MAX; ++q) { tv[q] = -INFINITY; ti[q] = 0x7fffffff; }
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) {
    const float v = to_f(xr[c]);
    mx = fmaxf(mx, v);
    if (v > tv[BK_MAX - 1]) {          // the list always keeps BK_MAX entries (static register indexing); k of them are used
      tv[BK_MAX - 1] = v; ti[BK_MAX - 1] = c;
#pragma unroll
      for (int q = BK_MAX - 1; q > 0; --q)
        if (tv[q] > tv[q - 1]) {
          const float fv = tv[q]; tv[q] = tv[q - 1]; tv[q - 1] = fv;
          const int fi = ti[q]; ti[q] = ti[q - 1]; ti[q - 1] = fi;
        }
    }
  }
  mx = block_max(mx, red);
  float se = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) se += expf(to_f(xr[c]) - mx);
  se = block_sum(se, red);
  const float lse = mx + logf(se);
  // warp-level merge: k rounds of "best head among the lanes", the winner pops its head
  for (int round = 0; round < k; ++round) {
    const float v = tv[0];
    const int i = ti[0];
    float bv = v; int bi = i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[w][round] = bv; si[w][round] = bi; }
    if (i == bi && bi != 0x7fffffff) {       // this lane won: shift its list up
#pragma unroll
      for (int q = 0; q < BK_MAX - 1; ++q) { tv[q] = tv[q + 1]; ti[q] = ti[q + 1]; }
      tv[BK_MAX - 1] = -INFINITY; ti[BK_MAX - 1] = 0x7fffffff;
    }
  }
  __syncthreads();
  // block-level merge of the 8 warp lists by warp 0 (lane = warp list index)
  if (w == 0) {
    int ptr = 0;
    for (int round = 0; round < k; ++round) {
      float v = (lane < 8 && ptr < k) ? sv[lane][ptr] : -INFINITY;
      int i = (lane < 8 && ptr < k) ? si[lane][ptr] : 0x7fffffff;
      float bv = v; int bi = i;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      bv = __shfl_sync(0xffffffffu, bv, 0);
      bi = __shfl_sync(0xffffffffu, bi, 0);
      if (lane < 8 && i == bi && bi != 0x7fffffff) ++ptr;
      if (lane == 0) {
        score[r * k + round] = bv - lse;
        pred[r * k + round] = (bi == 0x7fffffff) ? 0 : bi;
      }
    }
  }
}

// One CTA per utterance u; hypotheses u*k .. u*k+k-1.  `first` = decode position 1: the k beams of an utterance are identical,
// the first beam's top-k seeds them (Seq2seq.py:349-356).  eos / len_map are per SLOT and are not re-ordered (the reference
// does not re-order them either, Seq2seq.py:384-387).
__global__ void __launch_bounds__(256) beam_select_kernel(float* __restrict__ scores, const float* __restrict__ cand_score,
                                                          const int64_t* __restrict__ cand_pred, uint8_t* __restrict__ eos,
                                                          float* __restrict__ len_map, float penalty, int pos, int first,
                                                          int64_t* __restrict__ preds, int64_t ld_preds,
                                                          int32_t* __restrict__ anc, uint8_t* __restrict__ tokmask,
                                                          int64_t ld_tok, int k, int n_hyp, int32_t* __restrict__ done_u,
                                                          unsigned int* __restrict__ ticket, int64_t* __restrict__ n_done) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ int s_src[BK_MAX];
  __shared__ int64_t s_tok[BK_MAX];
  __shared__ int s_last;
  const int u = blockIdx.x, h0 = u * k, lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    // len^alpha (x ** 1 is x exactly, like torch.pow)
    auto lpow = [&](int slot) { const float l = len_map[h0 + slot]; return penalty == 1.f ? l : powf(l, penalty); };
    if (first) {
      if (lane < k) {
        scores[h0 + lane] += cand_score[(int64_t)h0 * k + lane];
        s_tok[lane] = cand_pred[(int64_t)h0 * k + lane];
        s_src[lane] = lane;
      }
    } else {
      // candidates c = r * k + j held two per lane (k * k <= 64)
      float v[2]; int id[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = lane + 32 * q;
        v[q] = -INFINITY; id[q] = 0x7fffffff;
        if (c < k * k) {
          const int rr = c / k, j = c % k;
          const bool done = eos[h0 + rr] != 0;
          const float sc = done ? (j == 0 ? 0.f : -1e9f) : cand_score[(int64_t)(h0 + rr) * k + j];      // Seq2seq.py:361-365
          v[q] = (scores[h0 + rr] + sc) / lpow(rr);                                                      // Seq2seq.py:367-371
          id[q] = c;
        }
      }
      float newscore = 0.f;
      for (int m = 0; m < k; ++m) {
        float bv = v[0]; int bi = id[0];
        if (v[1] > bv || (v[1] == bv && id[1] < bi)) { bv = v[1]; bi = id[1]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) if (id[q] == bi) { v[q] = -INFINITY; id[q] = 0x7fffffff; }
        if (lane == m) {
          newscore = bv * lpow(m);                                                                       // Seq2seq.py:373
          s_src[m] = bi / k;
          s_tok[m] = cand_pred[(int64_t)(h0 + bi / k) * k + bi % k];
        }
      }
      __syncwarp();
      if (lane < k) scores[h0 + lane] = newscore;
    }
  }
  __syncthreads();
  if (!first) {
    // re-order the prefixes of this utterance's hypotheses (Seq2seq.py:381-383) and the KV-cache bookkeeping: position p of
    // every slot is handled by ONE thread (all reads before its writes), so the permutation needs no staging
    for (int p = threadIdx.x; p < pos; p += 256) {
      int64_t t[BK_MAX]; int32_t a[BK_MAX]; uint8_t mk[BK_MAX];
#pragma unroll
      for (int m = 0; m < BK_MAX; ++m)
        if (m < k) {
          const int src = h0 + s_src[m];
          t[m] = preds[(int64_t)src * ld_preds + p];
          if (anc) a[m] = anc[(int64_t)p * n_hyp + src];
          mk[m] = tokmask[(int64_t)src * ld_tok + p];
        }
#pragma unroll
      for (int m = 0; m < BK_MAX; ++m)
        if (m < k) {
          preds[(int64_t)(h0 + m) * ld_preds + p] = t[m];
          if (anc) anc[(int64_t)p * n_hyp + h0 + m] = a[m];
          tokmask[(int64_t)(h0 + m) * ld_tok + p] = mk[m];
        }
    }
  }
  if (threadIdx.x == 0) {
    int done = 0;
    for (int m = 0; m < k; ++m) {
      const int64_t tok = s_tok[m];
      preds[(int64_t)(h0 + m) * ld_preds + pos] = tok;
      const bool e = (tok == 3 /*EOS*/) || eos[h0 + m] != 0;                                           // Seq2seq.py:384-385
      eos[h0 + m] = e ? 1 : 0;
      if (!e) len_map[h0 + m] += 1.f;                                                                  // Seq2seq.py:386-387
      done += e ? 1 : 0;
    }
    done_u[u] = done;
    __threadfence();
    s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last) {          // the last utterance to finish sums the per-utterance counts (the reference's eos_mask.sum())
    __shared__ int acc[256];
    int s = 0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) s += __ldcg(done_u + i);
    acc[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) acc[threadIdx.x] += acc[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) { *n_done = acc[0]; *ticket = 0u; }
  }
}

}  // namespace b200st

using namespace b200st;

extern "C" {

int b200st_topk_logsoftmax(int dtype, const void* x, int64_t ld, int64_t rows, int64_t cols, int64_t k, float* score,
                           int64_t* pred, b200st_stream_t stream) {
  if (rows <= 0) return 0;
  if (k < 1 || k > BK_MAX || k > cols) return set_error("topk_logsoftmax: k %lld out of range [1, %d]", (long long)k, BK_MAX);
  B200ST_DISPATCH(dtype, T, {
    B200ST_CUDA(launch_pdl(topk_logsoftmax_kernel<T>, dim3((unsigned)rows), dim3(256), 0, (cudaStream_t)stream, (const T*)x, ld,
                           (int)cols, (int)k, score, pred));
  });
  B200ST_LAUNCH_CHECK("topk_logsoftmax");
  return 0;
}

int b200st_beam_select(float* scores, const float* cand_score, const int64_t* cand_pred, uint8_t* eos, float* len_map,
                       float penalty, int64_t pos, int first, int64_t* preds, int64_t ld_preds, int32_t* anc,
                       uint8_t* tokmask, int64_t ld_tok, int64_t k, int64_t n_utt, int32_t* done_u, void* ticket,
                       int64_t* n_done, b200st_stream_t stream) {
  if (n_utt <= 0) return 0;
  if (k < 1 || k > BK_MAX) return set_error("beam_select: beam width %lld out of range [1, %d]", (long long)k, BK_MAX);
  B200ST_CUDA(launch_pdl(beam_select_kernel, dim3((unsigned)n_utt), dim3(256), 0, (cudaStream_t)stream, scores, cand_score,
                         cand_pred, eos, len_map, penalty, (int)pos, first, preds, ld_preds, anc, tokmask, ld_tok, (int)k,
                         (int)(n_utt * k), done_u, (unsigned int*)ticket, n_done));
  B200ST_LAUNCH_CHECK("beam_select");
  return 0;
}

}  // extern "C"
