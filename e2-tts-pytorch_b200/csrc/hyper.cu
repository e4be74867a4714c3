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

template <int VPT>
__device__ __forceinline__ void token_forward(const HcP& p, long long tok, int lane, TokState<VPT>& st) {
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
    const float sqrtD = sqrtf((float)p.D);
#pragma unroll
    for (int s = 0; s < HS; ++s) st.inv[s] = sqrtD / fmaxf(sqrtf(warp_sum(ss[s])), 1e-12f);

    float wc[HS][HT], dc[HS];
#pragma unroll
    for (int s = 0; s < HS; ++s) {
        dc[s] = 0.f;
#pragma unroll
        for (int t = 0; t < HT; ++t) wc[s][t] = 0.f;
    }
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        const int c = lane + 32 * v;
        if (c < nchunk) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = c * 8 + e;
                const float g1 = __ldg(p.gamma + i) + 1.f;
                const float bf = __ldg(p.bfn + i);
                float af[HT];
#pragma unroll
                for (int t = 0; t < HT; ++t) af[t] = __ldg(p.afn + i * HT + t);
#pragma unroll
                for (int s = 0; s < HS; ++s) {
                    const float nh = st.r[s][v][e] * st.inv[s] * g1;
                    dc[s] += nh * bf;
#pragma unroll
                    for (int t = 0; t < HT; ++t) wc[s][t] += nh * af[t];
                }
            }
        }
    }
    const float sa = __ldg(p.ascale), sb = __ldg(p.bscale);
#pragma unroll
    for (int s = 0; s < HS; ++s) {
        st.thb[s] = tanhf(warp_sum(dc[s]));
        st.beta[s] = st.thb[s] * sb + __ldg(p.sbeta + s);
#pragma unroll
        for (int t = 0; t < HT; ++t) {
            st.tha[s][t] = tanhf(warp_sum(wc[s][t]));
            st.alpha[s][t] = st.tha[s][t] * sa + __ldg(p.salpha + s * HT + t);
        }
    }
}

__device__ __forceinline__ const float* norm_gain(const HcP& p, long long tok) {
    return p.norm_mode == 2 ? p.ng + (size_t)(tok / p.rows_per_batch) * p.D : p.ng;
}

template <int VPT>
__global__ void __launch_bounds__(256) hc_width_fwd_kernel(const HcP p) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * 8;
    const int nchunk = p.D >> 3;
    for (long long tok = warp_global; tok < p.T; tok += nwarps) {
        TokState<VPT> st;
        token_forward<VPT>(p, tok, lane, st);
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
        if (lane < HS) p.beta_out[(size_t)tok * HS + lane] = st.beta[lane];
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

// Backward. grid = (ceil(rows_per_batch / 64), B): a block never straddles two batch elements, so the
// adaptive-gain gradient (B, D) can be accumulated in shared memory and flushed once per block.
constexpr int HC_TOK_PER_BLOCK = 64;

template <int VPT>
__global__ void __launch_bounds__(256) hc_width_bwd_kernel(const HcP p) {
    extern __shared__ float sacc[];  // [8][D] : afn t0..t4, bfn, gamma, ng   then [32] scalars
    const int D = p.D, nchunk = D >> 3;
    float* s_scal = sacc + 8 * D;
    for (int i = threadIdx.x; i < 8 * D + 32; i += 256) sacc[i] = 0.f;
    __syncthreads();
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
        token_forward<VPT>(p, tok, lane, st);

        // ---- branch (mix_0), its norm, and d(mix_0)
        float dm0[VPT][8];
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
                const float cn = sqrtf((float)D) / fmaxf(sqrtf(warp_sum(bss)), 1e-12f);
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
                    if (c < nchunk) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int i = c * 8 + e;
                            atomicAdd(&sacc[7 * D + i], dy[v][e] * br[v][e] * cn);
                            dm0[v][e] = cn * __ldg(ng + i) * dy[v][e] - br[v][e] * k2;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) dm0[v][e] = 0.f;
                    }
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
        float dal[HS][HT];
#pragma unroll
        for (int s = 0; s < HS; ++s)
#pragma unroll
            for (int t = 0; t < HT; ++t) dal[s][t] = 0.f;
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            const int c = lane + 32 * v;
#pragma unroll
            for (int s = 0; s < HS; ++s)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    dr[s][v][e] = st.alpha[s][0] * dm0[v][e];
                    dal[s][0] += dm0[v][e] * st.r[s][v][e];
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
                            dal[s][t] += dm[e] * st.r[s][v][e];
                        }
                }
            }
        }
        float dwc[HS][HT], ddc[HS];
#pragma unroll
        for (int s = 0; s < HS; ++s) {
            const float dbe = p.d_beta ? __ldg(p.d_beta + (size_t)tok * HS + s) : 0.f;
            ddc[s] = dbe * sb * (1.f - st.thb[s] * st.thb[s]);
            g_bs += dbe * st.thb[s];
            g_sbe[s] += dbe;
#pragma unroll
            for (int t = 0; t < HT; ++t) {
                const float da = warp_sum(dal[s][t]);
                dwc[s][t] = da * sa * (1.f - st.tha[s][t] * st.tha[s][t]);
                g_as += da * st.tha[s][t];
                g_sal[s][t] += da;
            }
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
                    const int i = c * 8 + e;
                    const float g1 = __ldg(p.gamma + i) + 1.f;
                    const float bf = __ldg(p.bfn + i);
                    float af[HT], gaf[HT];
#pragma unroll
                    for (int t = 0; t < HT; ++t) { af[t] = __ldg(p.afn + i * HT + t); gaf[t] = 0.f; }
                    float gbf = 0.f, ggam = 0.f;
#pragma unroll
                    for (int s = 0; s < HS; ++s) {
                        const float rn = st.r[s][v][e] * st.inv[s];
                        const float nh = rn * g1;
                        float dnh = ddc[s] * bf;
#pragma unroll
                        for (int t = 0; t < HT; ++t) { dnh += dwc[s][t] * af[t]; gaf[t] += nh * dwc[s][t]; }
                        gbf += nh * ddc[s];
                        ggam += dnh * rn;
                        const float uu = dnh * g1;
                        u[s][v][e] = uu;
                        R[s] += uu * st.r[s][v][e];
                    }
#pragma unroll
                    for (int t = 0; t < HT; ++t) atomicAdd(&sacc[t * D + i], gaf[t]);
                    atomicAdd(&sacc[5 * D + i], gbf);
                    atomicAdd(&sacc[6 * D + i], ggam);
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
    for (int i = threadIdx.x; i < D; i += 256) {
#pragma unroll
        for (int t = 0; t < HT; ++t) atomicAdd(p.g_afn + i * HT + t, sacc[t * D + i]);
        atomicAdd(p.g_bfn + i, sacc[5 * D + i]);
        atomicAdd(p.g_gamma + i, sacc[6 * D + i]);
        if (p.norm_mode == 1) atomicAdd(p.g_ng + i, sacc[7 * D + i]);
        else if (p.norm_mode == 2) atomicAdd(p.g_ng + (size_t)b * D + i, sacc[7 * D + i]);
    }
    if (threadIdx.x < 20) atomicAdd(p.g_salpha + threadIdx.x, s_scal[threadIdx.x]);
    else if (threadIdx.x < 24) atomicAdd(p.g_sbeta + (threadIdx.x - 20), s_scal[threadIdx.x]);
    else if (threadIdx.x == 24) atomicAdd(p.g_ascale, s_scal[24]);
    else if (threadIdx.x == 25) atomicAdd(p.g_bscale, s_scal[25]);
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
    if (a->D <= 256) hc_width_fwd_kernel<1><<<grid, 256, 0, st>>>(p);
    else if (a->D <= 512) hc_width_fwd_kernel<2><<<grid, 256, 0, st>>>(p);
    else hc_width_fwd_kernel<4><<<grid, 256, 0, st>>>(p);
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
    dim3 grid((a->rows_per_batch + HC_TOK_PER_BLOCK - 1) / HC_TOK_PER_BLOCK, a->T / a->rows_per_batch);
    const size_t smem = (size_t)(8 * a->D + 32) * sizeof(float);
    if (a->D <= 256) hc_width_bwd_kernel<1><<<grid, 256, smem, st>>>(p);
    else if (a->D <= 512) hc_width_bwd_kernel<2><<<grid, 256, smem, st>>>(p);
    else hc_width_bwd_kernel<4><<<grid, 256, smem, st>>>(p);
    return check_launch("hc_width_bwd_kernel");
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
