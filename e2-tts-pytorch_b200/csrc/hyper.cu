// Hyper-connection residual-stream kernels (S = 4 streams), fused with the consumer's (adaptive) RMSNorm.
//
// Replaces hyper_connections.HyperConnections as composed by the reference (ctor e2_tts.py:607,
// 673-678, 709-713; calls :870-882, :900-939) together with the x-transformers RMSNorm / AdaptiveRMSNorm that
// consumes the branch input (A.1; :875, :881, :908, :937):
//   width : n^ = RMSNorm_{gamma+1}(r_s);  alpha = tanh(n^ A) * sa + alpha0;  beta = tanh(n^ b) * sb + beta0
//           mix_t = sum_s alpha[s,t] r_s;  branch = mix_0 (optionally normalised);  residual'_t = mix_{t+1}
//   depth : out_s = residual'_s + beta_s * y
// HBM layout: residual streams are (token, stream, d) bf16 so the 4 streams of a token are adjacent (the
// reference's '(b s) n d' puts them N'*d apart). One warp owns one token; all reductions are warp shuffles.
// These kernels are HBM-bound: width reads S*d and writes (S+1)*d bf16 per token (algorithmic minimum).
#include "common.cuh"
#include "ptx.cuh"

namespace b200 {

constexpr int HS = 4;         // residual streams (reference default num_residual_streams = 4, e2_tts.py:547)
constexpr int HT = HS + 1;

struct HcP {
    const __nv_bfloat16* xres;  // [T, S, D]
    const float *gamma, *afn, *ascale, *salpha, *bfn, *bscale, *sbeta;
    int norm_mode;              // 0 none, 1 RMSNorm gain g[D], 2 adaptive gain (1+gamma)[B, D]
    const float* ng;
    int rows_per_batch, T, D;
    __nv_bfloat16 *branch, *res_out;
    float* beta_out;
    // backward
    const __nv_bfloat16 *d_branch, *d_res;
    const float* d_beta;
    __nv_bfloat16* d_xres;
    float *g_gamma, *g_afn, *g_ascale, *g_salpha, *g_bfn, *g_bscale, *g_sbeta, *g_ng;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Reduce 32 per-lane values across the warp with recursive halving (31 shuffles instead of 160): on return
// lane l holds the warp-wide sum of v[l] in v[0].
template <int N>
__device__ __forceinline__ void warp_halve(float (&v)[32], int lane) {
    constexpr int o = N / 2;
    const bool hi = lane & o;
#pragma unroll
    for (int i = 0; i < o; ++i) {
        const float send = hi ? v[i] : v[i + o];
        const float keep = hi ? v[i + o] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
}
__device__ __forceinline__ float warp_reduce32(float (&v)[32], int lane) {
    warp_halve<32>(v, lane);
    warp_halve<16>(v, lane);
    warp_halve<8>(v, lane);
    warp_halve<4>(v, lane);
    warp_halve<2>(v, lane);
    return v[0];
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
    f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// Per-token forward state shared by the forward and backward kernels.
template <int VPT>
struct TokState {
    float r[HS][VPT][8];
    float inv[HS];        // sqrt(D) / max(||r_s||, 1e-12)
    float tha[HS][HT];    // tanh(n^_s . A[:,t])
    float thb[HS];        // tanh(n^_s . b)
    float alpha[HS][HT];
    float beta[HS];
};

// Stage the per-feature parameters once per block: sp[i] = { (gamma_i+1) * A[i][0..4], (gamma_i+1) * b[i], gamma_i+1, 0 }
// (two 16-byte shared loads per feature instead of seven global loads; the (gamma+1) factor is folded in).
// Layout: feature i = chunk*8 + e lives at sp[(e*2 + part) * nchunk + chunk], so the 32 lanes of a warp (consecutive chunks,
// same e) read 32 consecutive float4 — bank-conflict free (the naive [i][2] layout was an 8-way conflict, ncu r1).
__device__ __forceinline__ int sp_idx(int nchunk, int chunk, int e, int part) { return (e * 2 + part) * nchunk + chunk; }
__device__ __forceinline__ void stage_params(const HcP& p, float4* sp) {
    const int nchunk = p.D >> 3;
    for (int i = threadIdx.x; i < p.D; i += blockDim.x) {
        const float g1 = __ldg(p.gamma + i) + 1.f;
        const float* a = p.afn + i * HT;
        sp[sp_idx(nchunk, i >> 3, i & 7, 0)] = make_float4(g1 * __ldg(a), g1 * __ldg(a + 1), g1 * __ldg(a + 2), g1 * __ldg(a + 3));
        sp[sp_idx(nchunk, i >> 3, i & 7, 1)] = make_float4(g1 * __ldg(a + 4), g1 * __ldg(p.bfn + i), g1, 0.f);
    }
    __syncthreads();
}

template <int VPT>
__device__ __forceinline__ void token_forward(const HcP& p, const float4* __restrict__ sp, long long tok, int lane, TokState<VPT>& st) {
    const int nchunk = p.D >> 3;
    float ss[HS] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const int c = lane + 32 * v;
#pragma unroll
        for (int s = 0; s < HS; ++s) {
            if (c < nchunk) {
                const uint4 u = *reinterpret_cast<const uint4*>(p.xres + ((size_t)tok * HS + s) * p.D + c * 8);
                unpack8(u, st.r[s][v]);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) st.r[s][v][e] = 0.f;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) ss[s] += st.r[s][v][e] * st.r[s][v][e];
        }
    }
    float red[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) red[i] = 0.f;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const int c = lane + 32 * v;
        if (c < nchunk) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 p0 = sp[sp_idx(nchunk, c, e, 0)], p1 = sp[sp_idx(nchunk, c, e, 1)];
#pragma unroll
                for (int s = 0; s < HS; ++s) {
                    const float rv = st.r[s][v][e];
                    red[s * HT + 0] += rv * p0.x; red[s * HT + 1] += rv * p0.y; red[s * HT + 2] += rv * p0.z;
                    red[s * HT + 3] += rv * p0.w; red[s * HT + 4] += rv * p1.x; red[HS * HT + s] += rv * p1.y;
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < HS; ++s) red[28 + s] = ss[s];   // the four sums of squares ride along in the same reduction
    const float mine = warp_reduce32(red, lane);         // lane l owns total #l
    const float sqrtD = sqrtf((float)p.D);
    float invs[HS];
#pragma unroll
    for (int s = 0; s < HS; ++s) {
        invs[s] = sqrtD / fmaxf(sqrtf(__shfl_sync(0xffffffffu, mine, 28 + s)), 1e-12f);
        st.inv[s] = invs[s];
    }
    // each lane applies inv_s and tanh to the ONE dot product it owns, then the 24 results are broadcast
    float myinv = invs[0];
    {
        const int s_of = lane < HS * HT ? lane / HT : lane - HS * HT;
#pragma unroll
        for (int s = 1; s < HS; ++s) myinv = (s_of == s) ? invs[s] : myinv;
    }
    const float th = tanhf(mine * myinv);
    const float sa = __ldg(p.ascale), sb = __ldg(p.bscale);
#pragma unroll
    for (int s = 0; s < HS; ++s) {
        st.thb[s] = __shfl_sync(0xffffffffu, th, HS * HT + s);
        st.beta[s] = st.thb[s] * sb + __ldg(p.sbeta + s);
#pragma unroll
        for (int t = 0; t < HT; ++t) {
            st.tha[s][t] = __shfl_sync(0xffffffffu, th, s * HT + t);
            st.alpha[s][t] = st.tha[s][t] * sa + __ldg(p.salpha + s * HT + t);
        }
    }
}

__device__ __forceinline__ const float* norm_gain(const HcP& p, long long tok) {
    return p.norm_mode == 2 ? p.ng + (size_t)(tok / p.rows_per_batch) * p.D : p.ng;
}

template <int VPT>
__global__ void __launch_bounds__(256, (VPT <= 2) ? 2 : 1) hc_width_fwd_kernel(const HcP p) {
    extern __shared__ float4 sp[];
    stage_params(p, sp);
    const int lane = threadIdx.x & 31;
    const long long warp_global = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    const int nchunk = p.D >> 3;
    for (long long tok = warp_global; tok < p.T; tok += nwarps) {
        TokState<VPT> st;
        token_forward<VPT>(p, sp, tok, lane, st);
        float br[VPT][8];
        float bss = 0.f;
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            const int c = lane + 32 * v;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float acc = 0.f;
#pragma unroll
                for (int s = 0; s < HS; ++s) acc += st.alpha[s][0] * st.r[s][v][e];
                br[v][e] = acc;
                bss += acc * acc;
            }
            if (c < nchunk) {
#pragma unroll
                for (int t = 1; t < HT; ++t) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float acc = 0.f;
#pragma unroll
                        for (int s = 0; s < HS; ++s) acc += st.alpha[s][t] * st.r[s][v][e];
                        o[e] = acc;
                    }
                    *reinterpret_cast<uint4*>(p.res_out + ((size_t)tok * HS + (t - 1)) * p.D + c * 8) = pack8(o);
                }
            }
        }
        {
            float bsel = st.beta[0];
#pragma unroll
            for (int s = 1; s < HS; ++s) bsel = (lane == s) ? st.beta[s] : bsel;
            if (lane < HS) p.beta_out[(size_t)tok * HS + lane] = bsel;
        }
        float c_norm = 1.f;
        const float* ng = nullptr;
        if (p.norm_mode) {
            c_norm = sqrtf((float)p.D) / fmaxf(sqrtf(warp_sum(bss)), 1e-12f);
            ng = norm_gain(p, tok);
        }
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            const int c = lane + 32 * v;
            if (c < nchunk) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = p.norm_mode ? br[v][e] * c_norm * __ldg(ng + c * 8 + e) : br[v][e];
                *reinterpret_cast<uint4*>(p.branch + (size_t)tok * p.D + c * 8) = pack8(o);
            }
        }
    }
}

// Backward, two kernels so that no per-element atomics are needed:
//   (1) token kernel  : one warp per token — recompute the forward scalars, produce d_xres and a 40-float per-token
//                       record (inv[4], d_wc[20], d_dc[4], alpha[:,0][4], c_norm, pad) plus the scalar parameter grads;
//   (2) param kernel  : one thread per pair of feature columns, marching over a slab of tokens with the records in
//                       shared memory — accumulates d(dynamic_alpha_fn), d(dynamic_beta_fn), d(norm.gamma), d(gain)
//                       in registers; one global atomicAdd per column per block.
// grid.y = batch element: a block never straddles two batch elements (adaptive-gain gradient is per batch).
constexpr int HC_TOK_PER_BLOCK = 64;
constexpr int HC_REC = 40;
#ifndef HC_BWD_MIN_BLOCKS
#define HC_BWD_MIN_BLOCKS 1
#endif

template <int VPT>
__global__ void __launch_bounds__(256, HC_BWD_MIN_BLOCKS) hc_width_bwd_kernel(const HcP p, float* __restrict__ rec) {
    extern __shared__ float4 sp[];
    __shared__ float s_scal[32];
    if (threadIdx.x < 32) s_scal[threadIdx.x] = 0.f;
    stage_params(p, sp);
    const int D = p.D, nchunk = D >> 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * HC_TOK_PER_BLOCK;
    const int n1 = min(p.rows_per_batch, n0 + HC_TOK_PER_BLOCK);
    float g_sal[HS][HT], g_sbe[HS], g_as = 0.f, g_bs = 0.f;
#pragma unroll
    for (int s = 0; s < HS; ++s) {
        g_sbe[s] = 0.f;
#pragma unroll
        for (int t = 0; t < HT; ++t) g_sal[s][t] = 0.f;
    }
    const float sa = __ldg(p.ascale), sb = __ldg(p.bscale);
    const float invD = 1.f / (float)D;

    for (int n = n0 + warp; n < n1; n += 8) {
        const long long tok = (long long)b * p.rows_per_batch + n;
        TokState<VPT> st;
        token_forward<VPT>(p, sp, tok, lane, st);

        // ---- branch (mix_0), its norm, and d(mix_0)
        float dm0[VPT][8];
        float cn = 1.f;
        {
            float br[VPT][8], dy[VPT][8];
            float bss = 0.f;
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
                const int c = lane + 32 * v;
                if (c < nchunk) unpack8(*reinterpret_cast<const uint4*>(p.d_branch + (size_t)tok * D + c * 8), dy[v]);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (c >= nchunk) dy[v][e] = 0.f;
                    float acc = 0.f;
#pragma unroll
                    for (int s = 0; s < HS; ++s) acc += st.alpha[s][0] * st.r[s][v][e];
                    br[v][e] = acc;
                    bss += acc * acc;
                }
            }
            if (p.norm_mode) {
                cn = sqrtf((float)D) / fmaxf(sqrtf(warp_sum(bss)), 1e-12f);
                const float* ng = norm_gain(p, tok);
                float dot = 0.f;
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    const int c = lane + 32 * v;
                    if (c < nchunk) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) dot += __ldg(ng + c * 8 + e) * dy[v][e] * br[v][e];
                    }
                }
                dot = warp_sum(dot);
                const float k2 = cn * cn * cn * invD * dot;
#pragma unroll
                for (int v = 0; v < VPT; ++v) {
                    const int c = lane + 32 * v;
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        dm0[v][e] = (c < nchunk) ? cn * __ldg(ng + c * 8 + e) * dy[v][e] - br[v][e] * k2 : 0.f;
                }
            } else {
#pragma unroll
                for (int v = 0; v < VPT; ++v)
#pragma unroll
                    for (int e = 0; e < 8; ++e) dm0[v][e] = dy[v][e];
            }
        }
        // ---- d_alpha[s][t] = <d_mix_t, r_s>; start d_r_s = sum_t alpha[s][t] d_mix_t
        float dr[HS][VPT][8];
        float red[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) red[i] = 0.f;
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            const int c = lane + 32 * v;
#pragma unroll
            for (int s = 0; s < HS; ++s)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    dr[s][v][e] = st.alpha[s][0] * dm0[v][e];
                    red[s * HT] += dm0[v][e] * st.r[s][v][e];
                }
            if (c < nchunk) {
#pragma unroll
                for (int t = 1; t < HT; ++t) {
                    float dm[8];
                    unpack8(*reinterpret_cast<const uint4*>(p.d_res + ((size_t)tok * HS + (t - 1)) * D + c * 8), dm);
#pragma unroll
                    for (int s = 0; s < HS; ++s)
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            dr[s][v][e] += st.alpha[s][t] * dm[e];
                            red[s * HT + t] += dm[e] * st.r[s][v][e];
                        }
                }
            }
        }
        const float mine = warp_reduce32(red, lane);
        float dwc[HS][HT], ddc[HS];
#pragma unroll
        for (int s = 0; s < HS; ++s) {
            const float dbe = p.d_beta ? __ldg(p.d_beta + (size_t)tok * HS + s) : 0.f;
            ddc[s] = dbe * sb * (1.f - st.thb[s] * st.thb[s]);
            g_bs += dbe * st.thb[s];
            g_sbe[s] += dbe;
#pragma unroll
            for (int t = 0; t < HT; ++t) {
                const float da = __shfl_sync(0xffffffffu, mine, s * HT + t);
                dwc[s][t] = da * sa * (1.f - st.tha[s][t] * st.tha[s][t]);
                g_as += da * st.tha[s][t];
                g_sal[s][t] += da;
            }
        }
        // per-token record for the parameter kernel
        {
            float val = 0.f;
#pragma unroll
            for (int s = 0; s < HS; ++s) {
                if (lane == s) val = st.inv[s];
                if (lane == 24 + s) val = ddc[s];
                if (lane == 28 + s) val = st.alpha[s][0];
#pragma unroll
                for (int t = 0; t < HT; ++t)
                    if (lane == 4 + s * HT + t) val = dwc[s][t];
            }
            rec[(size_t)tok * HC_REC + lane] = val;
            if (lane == 0) rec[(size_t)tok * HC_REC + 32] = cn;
        }
        // ---- through n^ = r * inv * (gamma+1)
        float R[HS] = {0.f, 0.f, 0.f, 0.f};
        float u[HS][VPT][8];
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            const int c = lane + 32 * v;
            if (c < nchunk) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float4 p0 = sp[sp_idx(nchunk, c, e, 0)], p1 = sp[sp_idx(nchunk, c, e, 1)];   // already scaled by (gamma+1)
#pragma unroll
                    for (int s = 0; s < HS; ++s) {
                        const float uu = ddc[s] * p1.y + dwc[s][0] * p0.x + dwc[s][1] * p0.y + dwc[s][2] * p0.z + dwc[s][3] * p0.w + dwc[s][4] * p1.x;
                        u[s][v][e] = uu;
                        R[s] += uu * st.r[s][v][e];
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < HS; ++s)
#pragma unroll
                    for (int e = 0; e < 8; ++e) u[s][v][e] = 0.f;
            }
        }
#pragma unroll
        for (int s = 0; s < HS; ++s) {
            const float Rs = warp_sum(R[s]);
            const float k3 = st.inv[s] * st.inv[s] * st.inv[s] * invD * Rs;
#pragma unroll
            for (int v = 0; v < VPT; ++v) {
                const int c = lane + 32 * v;
                if (c < nchunk) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = dr[s][v][e] + st.inv[s] * u[s][v][e] - st.r[s][v][e] * k3;
                    *reinterpret_cast<uint4*>(p.d_xres + ((size_t)tok * HS + s) * D + c * 8) = pack8(o);
                }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < HS; ++s) {
            atomicAdd(&s_scal[20 + s], g_sbe[s]);
#pragma unroll
            for (int t = 0; t < HT; ++t) atomicAdd(&s_scal[s * HT + t], g_sal[s][t]);
        }
        atomicAdd(&s_scal[24], g_as);
        atomicAdd(&s_scal[25], g_bs);
    }
    __syncthreads();
    if (threadIdx.x < 20) atomicAdd(p.g_salpha + threadIdx.x, s_scal[threadIdx.x]);
    else if (threadIdx.x < 24) atomicAdd(p.g_sbeta + (threadIdx.x - 20), s_scal[threadIdx.x]);
    else if (threadIdx.x == 24) atomicAdd(p.g_ascale, s_scal[24]);
    else if (threadIdx.x == 25) atomicAdd(p.g_bscale, s_scal[25]);
}

constexpr int HC_PARAM_TOK = 128;   // tokens per block (4 groups of 32)
constexpr int HC_PARAM_COLS = 128;  // feature columns per block (64 threads x 2)

__global__ void __launch_bounds__(256) hc_width_bwd_param_kernel(const HcP p, const float* __restrict__ rec) {
    __shared__ float srec[HC_PARAM_TOK][36];
    __shared__ float sred[8][HC_PARAM_COLS];
    const int D = p.D, b = blockIdx.z;
    const int n0 = blockIdx.x * HC_PARAM_TOK;
    const int ntok = min(p.rows_per_batch - n0, HC_PARAM_TOK);
    const long long tok0 = (long long)b * p.rows_per_batch + n0;
    for (int i = threadIdx.x; i < ntok * 36; i += 256) srec[i / 36][i % 36] = rec[(size_t)(tok0 + i / 36) * HC_REC + (i % 36)];
    for (int i = threadIdx.x; i < 8 * HC_PARAM_COLS; i += 256) sred[i / HC_PARAM_COLS][i % HC_PARAM_COLS] = 0.f;
    __syncthreads();
    const int cp = threadIdx.x & 63, tg = threadIdx.x >> 6;
    const int col = blockIdx.y * HC_PARAM_COLS + cp * 2;
    if (col < D) {
        float g1[2], bf[2], af[2][HT], gaf[2][HT], gbf[2] = {0.f, 0.f}, ggam[2] = {0.f, 0.f}, gng[2] = {0.f, 0.f};
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            g1[q] = __ldg(p.gamma + col + q) + 1.f;
            bf[q] = __ldg(p.bfn + col + q);
#pragma unroll
            for (int t = 0; t < HT; ++t) { af[q][t] = __ldg(p.afn + (col + q) * HT + t); gaf[q][t] = 0.f; }
        }
        const int nbeg = tg * 32, nend = min(ntok, nbeg + 32);
#pragma unroll 4
        for (int n = nbeg; n < nend; ++n) {
            const float* rc = srec[n];
            const size_t tok = (size_t)(tok0 + n);
            float r[HS][2];
#pragma unroll
            for (int s = 0; s < HS; ++s) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(p.xres + (tok * HS + s) * D + col);
                r[s][0] = bf16_lo(w); r[s][1] = bf16_hi(w);
            }
            float dy[2] = {0.f, 0.f};
            if (p.norm_mode) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(p.d_branch + tok * D + col);
                dy[0] = bf16_lo(w); dy[1] = bf16_hi(w);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float bmix = 0.f;
#pragma unroll
                for (int s = 0; s < HS; ++s) {
                    const float rn = r[s][q] * rc[s];
                    const float nh = rn * g1[q];
                    const float ddc = rc[24 + s];
                    float dnh = ddc * bf[q];
#pragma unroll
                    for (int t = 0; t < HT; ++t) {
                        const float dwc = rc[4 + s * HT + t];
                        dnh += dwc * af[q][t];
                        gaf[q][t] += nh * dwc;
                    }
                    gbf[q] += nh * ddc;
                    ggam[q] += dnh * rn;
                    bmix += rc[28 + s] * r[s][q];
                }
                gng[q] += dy[q] * bmix * rc[32];
            }
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int lc = cp * 2 + q;
#pragma unroll
            for (int t = 0; t < HT; ++t) atomicAdd(&sred[t][lc], gaf[q][t]);
            atomicAdd(&sred[5][lc], gbf[q]);
            atomicAdd(&sred[6][lc], ggam[q]);
            atomicAdd(&sred[7][lc], gng[q]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * HC_PARAM_COLS; i += 256) {
        const int k = i / HC_PARAM_COLS, c = blockIdx.y * HC_PARAM_COLS + (i % HC_PARAM_COLS);
        if (c >= D) continue;
        const float v = sred[k][i % HC_PARAM_COLS];
        if (k < HT) atomicAdd(p.g_afn + c * HT + k, v);
        else if (k == 5) atomicAdd(p.g_bfn + c, v);
        else if (k == 6) atomicAdd(p.g_gamma + c, v);
        else if (p.norm_mode == 1) atomicAdd(p.g_ng + c, v);
        else if (p.norm_mode == 2) atomicAdd(p.g_ng + (size_t)b * D + c, v);
    }
}

// ------------------------------------------------------------------------------------------------ depth
struct HdP {
    const __nv_bfloat16 *res, *y;
    const float* beta;
    __nv_bfloat16* out;
    int T, D;
    const __nv_bfloat16* d_out;
    __nv_bfloat16* d_y;
    float* d_beta;
};

// out[t,s,:] = res[t,s,:] + beta[t,s] * y[t,:]     (one 16-byte chunk of y per thread, all 4 streams)
__global__ void __launch_bounds__(256) hc_depth_fwd_kernel(const HdP p) {
    const int nchunk = p.D >> 3;
    const long long total = (long long)p.T * nchunk;
    for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (long long)gridDim.x * 256) {
        const long long tok = idx / nchunk;
        const int c = (int)(idx % nchunk);
        float y[8];
        unpack8(*reinterpret_cast<const uint4*>(p.y + (size_t)tok * p.D + c * 8), y);
        const float4 be = *reinterpret_cast<const float4*>(p.beta + (size_t)tok * HS);
        const float bes[HS] = {be.x, be.y, be.z, be.w};
#pragma unroll
        for (int s = 0; s < HS; ++s) {
            float r[8];
            const size_t off = ((size_t)tok * HS + s) * p.D + c * 8;
            unpack8(*reinterpret_cast<const uint4*>(p.res + off), r);
#pragma unroll
            for (int e = 0; e < 8; ++e) r[e] += bes[s] * y[e];
            *reinterpret_cast<uint4*>(p.out + off) = pack8(r);
        }
    }
}

// d_y[t,:] = sum_s beta[t,s] d_out[t,s,:];  d_beta[t,s] = <d_out[t,s,:], y[t,:]>   (one warp per token)
__global__ void __launch_bounds__(256) hc_depth_bwd_kernel(const HdP p) {
    const int lane = threadIdx.x & 31;
    const int nchunk = p.D >> 3;
    const long long nwarps = (long long)gridDim.x * 8;
    for (long long tok = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); tok < p.T; tok += nwarps) {
        const float4 be = *reinterpret_cast<const float4*>(p.beta + (size_t)tok * HS);
        const float bes[HS] = {be.x, be.y, be.z, be.w};
        float db[HS] = {0.f, 0.f, 0.f, 0.f};
        for (int c = lane; c < nchunk; c += 32) {
            float y[8], dy[8];
            unpack8(*reinterpret_cast<const uint4*>(p.y + (size_t)tok * p.D + c * 8), y);
#pragma unroll
            for (int e = 0; e < 8; ++e) dy[e] = 0.f;
#pragma unroll
            for (int s = 0; s < HS; ++s) {
                float d[8];
                unpack8(*reinterpret_cast<const uint4*>(p.d_out + ((size_t)tok * HS + s) * p.D + c * 8), d);
#pragma unroll
                for (int e = 0; e < 8; ++e) { dy[e] += bes[s] * d[e]; db[s] += d[e] * y[e]; }
            }
            *reinterpret_cast<uint4*>(p.d_y + (size_t)tok * p.D + c * 8) = pack8(dy);
        }
#pragma unroll
        for (int s = 0; s < HS; ++s) db[s] = warp_sum(db[s]);
        if (lane == 0) *reinterpret_cast<float4*>(p.d_beta + (size_t)tok * HS) = make_float4(db[0], db[1], db[2], db[3]);
    }
}

static int fill_hc(HcP& p, const b200_hc_width_args* a) {
    B200_REQUIRE(a->num_streams == HS, "hyper-connections: only num_residual_streams=4 is built (got %d)", a->num_streams);
    B200_REQUIRE(a->D >= 8 && (a->D % 8) == 0 && a->D <= 1024, "hyper-connections: D=%d must be a multiple of 8 and <= 1024", a->D);
    B200_REQUIRE(a->T > 0 && a->rows_per_batch > 0 && (a->T % a->rows_per_batch) == 0, "hyper-connections: T must be a multiple of rows_per_batch");
    B200_REQUIRE(a->norm_mode >= 0 && a->norm_mode <= 2 && (a->norm_mode == 0 || a->norm_gain), "hyper-connections: bad norm mode");
    p.xres = (const __nv_bfloat16*)a->xres;
    p.gamma = a->norm_gamma; p.afn = a->dynamic_alpha_fn; p.ascale = a->dynamic_alpha_scale; p.salpha = a->static_alpha;
    p.bfn = a->dynamic_beta_fn; p.bscale = a->dynamic_beta_scale; p.sbeta = a->static_beta;
    p.norm_mode = a->norm_mode; p.ng = a->norm_gain; p.rows_per_batch = a->rows_per_batch; p.T = a->T; p.D = a->D;
    return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_hc_width_fwd(const b200_hc_width_args* a, b200_stream_t stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B200_REQUIRE(a && a->xres && a->branch && a->res_out && a->beta_out, "hc_width_fwd: null pointer");
    HcP p{};
    if (fill_hc(p, a)) return -1;
    p.branch = (__nv_bfloat16*)a->branch; p.res_out = (__nv_bfloat16*)a->res_out; p.beta_out = a->beta_out;
    const int grid = (int)min((long long)(a->T + 7) / 8, (long long)num_sms() * 8);
    const size_t smem = (size_t)a->D * 2 * sizeof(float4);
    if (a->D <= 256) hc_width_fwd_kernel<1><<<grid, 256, smem, st>>>(p);
    else if (a->D <= 512) hc_width_fwd_kernel<2><<<grid, 256, smem, st>>>(p);
    else hc_width_fwd_kernel<4><<<grid, 256, smem, st>>>(p);
    return check_launch("hc_width_fwd_kernel");
}

extern "C" int b200_hc_width_bwd(const b200_hc_width_args* a, b200_stream_t stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B200_REQUIRE(a && a->xres && a->d_branch && a->d_res && a->d_xres && a->g_norm_gamma && a->g_dynamic_alpha_fn && a->g_dynamic_alpha_scale &&
                 a->g_static_alpha && a->g_dynamic_beta_fn && a->g_dynamic_beta_scale && a->g_static_beta, "hc_width_bwd: null pointer");
    HcP p{};
    if (fill_hc(p, a)) return -1;
    B200_REQUIRE(a->norm_mode == 0 || a->g_norm_gain, "hc_width_bwd: missing gain gradient buffer");
    p.d_branch = (const __nv_bfloat16*)a->d_branch; p.d_res = (const __nv_bfloat16*)a->d_res; p.d_beta = a->d_beta;
    p.d_xres = (__nv_bfloat16*)a->d_xres;
    p.g_gamma = a->g_norm_gamma; p.g_afn = a->g_dynamic_alpha_fn; p.g_ascale = a->g_dynamic_alpha_scale; p.g_salpha = a->g_static_alpha;
    p.g_bfn = a->g_dynamic_beta_fn; p.g_bscale = a->g_dynamic_beta_scale; p.g_sbeta = a->g_static_beta; p.g_ng = a->g_norm_gain;
    B200_REQUIRE(a->ws_records, "hc_width_bwd: missing per-token record workspace (T * 40 floats)");
    dim3 grid((a->rows_per_batch + HC_TOK_PER_BLOCK - 1) / HC_TOK_PER_BLOCK, a->T / a->rows_per_batch);
    const size_t smem = (size_t)a->D * 2 * sizeof(float4);
    if (a->D <= 256) hc_width_bwd_kernel<1><<<grid, 256, smem, st>>>(p, a->ws_records);
    else if (a->D <= 512) hc_width_bwd_kernel<2><<<grid, 256, smem, st>>>(p, a->ws_records);
    else hc_width_bwd_kernel<4><<<grid, 256, smem, st>>>(p, a->ws_records);
    if (int rc = check_launch("hc_width_bwd_kernel")) return rc;
    dim3 grid2((a->rows_per_batch + HC_PARAM_TOK - 1) / HC_PARAM_TOK, (a->D + HC_PARAM_COLS - 1) / HC_PARAM_COLS, a->T / a->rows_per_batch);
    hc_width_bwd_param_kernel<<<grid2, 256, 0, st>>>(p, a->ws_records);
    return check_launch("hc_width_bwd_param_kernel");
}

extern "C" int b200_hc_depth_fwd(const b200_hc_depth_args* a, b200_stream_t stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B200_REQUIRE(a && a->res && a->y && a->beta && a->out, "hc_depth_fwd: null pointer");
    B200_REQUIRE(a->num_streams == HS && a->D % 8 == 0 && a->T > 0, "hc_depth_fwd: unsupported shape");
    HdP p{};
    p.res = (const __nv_bfloat16*)a->res; p.y = (const __nv_bfloat16*)a->y; p.beta = a->beta; p.out = (__nv_bfloat16*)a->out; p.T = a->T; p.D = a->D;
    const long long total = (long long)a->T * (a->D / 8);
    const int grid = (int)min((total + 255) / 256, (long long)num_sms() * 16);
    hc_depth_fwd_kernel<<<grid, 256, 0, st>>>(p);
    return check_launch("hc_depth_fwd_kernel");
}

extern "C" int b200_hc_depth_bwd(const b200_hc_depth_args* a, b200_stream_t stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    B200_REQUIRE(a && a->y && a->beta && a->d_out && a->d_y && a->d_beta, "hc_depth_bwd: null pointer");
    B200_REQUIRE(a->num_streams == HS && a->D % 8 == 0 && a->T > 0, "hc_depth_bwd: unsupported shape");
    HdP p{};
    p.y = (const __nv_bfloat16*)a->y; p.beta = a->beta; p.T = a->T; p.D = a->D;
    p.d_out = (const __nv_bfloat16*)a->d_out; p.d_y = (__nv_bfloat16*)a->d_y; p.d_beta = a->d_beta;
    const int grid = (int)min((long long)(a->T + 7) / 8, (long long)num_sms() * 8);
    hc_depth_bwd_kernel<<<grid, 256, 0, st>>>(p);
    return check_launch("hc_depth_bwd_kernel");
}
