// Parallel-WaveNet student (4-flow IAF) forward for B200.
//
// Replaces ParallelWavenet.feed_forward / _create_iaf / _clip_quant_scale
// (wavenet/parallel_wavenet.py:200-359) as driven by parallelgen.load_parallelgen
// (wavenet/parallelgen.py:11-19).
//
// Data layout in HBM (all fp32, channels-last like the reference, masked.py:43):
//   l[2]      [B*T, 64]            residual stream, ping-pong (a layer reads t-d, t-2d of
//                                  other tiles, so it cannot update in place)
//   cond      [(L+1)][B*T, 64]     per-flow mel conditioning planes, biases folded in;
//                                  layer planes are stored gate-interleaved:
//                                  position 2j <- channel j (sigmoid half),
//                                  position 2j+1 <- channel j+32 (tanh half)
//   mel_en    [B, 200F, 256]       fp32 (FFMA engine) or fp16 hi + fp16 lo (tcgen05 engine)
//   x, z, mean_tot, scale_tot, log_scale_tot   [B*T]
//
// Kernels (one launch each):
//   iaf_start_conv_kernel   shift_right + start_conv (parallel_wavenet.py:222-225)
//   iaf_layer_kernel        dilated 3-tap conv + cond add + sigmoid*tanh gate + 1x1 res +
//                           residual add (parallel_wavenet.py:227-254), FFMA2
//   iaf_head_kernel         relu/out1/cond/relu/out2_mean/out2_scale, softplus-clip-log,
//                           x*s+m, running composition (:256-277, :316-330), final clip +
//                           quantise (:326-330, :348-359)
//   conv-GEMM               mel-cond projections for all layers of a flow, deconv stack
#include "nsw_gemm.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace nsw {

namespace {

constexpr int C = 64;    // residual width == gate width (parallel_wavenet.py:209)
constexpr int HALF = 32; // gate half
constexpr int D = 256;   // deconv_width

// ------------------------------- noise -------------------------------------
__global__ void iaf_noise_kernel(float* __restrict__ z, size_t n, uint64_t seed, int gauss) {
  const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)i4, (uint32_t)(i4 >> 32), 0x6e7377u, 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float v[4];
  if (gauss) {
    // Box-Muller, two pairs (parallel_wavenet.py:181-184)
    const float u0 = u01_clipped(r.x), u1 = u01_clipped(r.y);
    const float u2 = u01_clipped(r.z), u3 = u01_clipped(r.w);
    const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
    float s0, c0, s1, c1;
    sincosf(6.283185307179586f * u1, &s0, &c0);
    sincosf(6.283185307179586f * u3, &s1, &c1);
    v[0] = r0 * c0; v[1] = r0 * s0; v[2] = r1 * c1; v[3] = r1 * s1;
  } else {
    // logistic(0,1) = log u - log(1-u)  (parallel_wavenet.py:173-178)
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float u = u01_clipped(rr[j]);
      v[j] = logf(u) - logf(1.0f - u);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i4 * 4 + j < n) z[i4 * 4 + j] = v[j];
}

// ------------------------------ start conv ---------------------------------
// l0[t,c] = b[c] + W[0,c] x[t-3] + W[1,c] x[t-2] + W[2,c] x[t-1]
// (masked.shift_right masked.py:39-52 then conv1d k=3, d=1, Cin=1)
__global__ void __launch_bounds__(256)
iaf_start_conv_kernel(const float* __restrict__ x, float* __restrict__ l,
                      const float* __restrict__ w /*[3][64]*/, const float* __restrict__ b, int T,
                      size_t rows, __half* __restrict__ l_hi, __half* __restrict__ l_lo) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // (row, c4)
  const size_t row = idx >> 4;
  if (row >= rows) return;
  const int c = (int)(idx & 15) * 4;
  const int t = (int)(row % T);
  const float x1 = t >= 1 ? x[row - 1] : 0.0f;
  const float x2 = t >= 2 ? x[row - 2] : 0.0f;
  const float x3 = t >= 3 ? x[row - 3] : 0.0f;
  const float4 w0 = *reinterpret_cast<const float4*>(w + c);
  const float4 w1 = *reinterpret_cast<const float4*>(w + C + c);
  const float4 w2 = *reinterpret_cast<const float4*>(w + 2 * C + c);
  const float4 bb = *reinterpret_cast<const float4*>(b + c);
  float4 o;
  o.x = fmaf(w2.x, x1, fmaf(w1.x, x2, fmaf(w0.x, x3, bb.x)));
  o.y = fmaf(w2.y, x1, fmaf(w1.y, x2, fmaf(w0.y, x3, bb.y)));
  o.z = fmaf(w2.z, x1, fmaf(w1.z, x2, fmaf(w0.z, x3, bb.z)));
  o.w = fmaf(w2.w, x1, fmaf(w1.w, x2, fmaf(w0.w, x3, bb.w)));
  *reinterpret_cast<float4*>(l + row * C + c) = o;
  if (l_hi) {  // split-fp16 copy for the tcgen05 layer kernel's MMA operand
    const float f[4] = {o.x, o.y, o.z, o.w};
    __align__(8) __half hi[4], lo[4];
    uint32_t rmx = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      range_track(rmx, f[i]);
      hi[i] = __float2half_rn(f[i]);
      lo[i] = __float2half_rn(f[i] - __half2float(hi[i]));
    }
    range_commit(rmx);
    *reinterpret_cast<uint2*>(l_hi + row * C + c) = *reinterpret_cast<uint2*>(hi);
    *reinterpret_cast<uint2*>(l_lo + row * C + c) = *reinterpret_cast<uint2*>(lo);
  }
}

// fp32 view of a split (fp16 hi + fp16 lo) residual stream: the tcgen05 layer engine keeps l only
// as the split pair; the head kernel and the debug tap read this merged copy
__global__ void __launch_bounds__(256)
iaf_merge_split_kernel(const __half* __restrict__ hi, const __half* __restrict__ lo,
                       float* __restrict__ out, size_t n8) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 h = reinterpret_cast<const uint4*>(hi)[i], l = reinterpret_cast<const uint4*>(lo)[i];
  const __half2* hp = reinterpret_cast<const __half2*>(&h);
  const __half2* lp = reinterpret_cast<const __half2*>(&l);
  float o[8];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 a = __half22float2(hp[e]), b = __half22float2(lp[e]);
    o[2 * e] = a.x + b.x;
    o[2 * e + 1] = a.y + b.y;
  }
  reinterpret_cast<float4*>(out)[2 * i] = make_float4(o[0], o[1], o[2], o[3]);
  reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(o[4], o[5], o[6], o[7]);
}

// ------------------------------ residual layer ------------------------------
// One warp owns RT consecutive time steps; lane j owns the output-channel pair
// (j, j+32), i.e. both gate halves of gate channel j, so the gate is lane-local.
// Products run as FFMA2 with the activation as the broadcast scalar operand:
//   acc2[t] (+)= Wpair[k] * a[t][k]
// A tiles (the three taps t-2d, t-d, t) are staged with cp.async in natural layout
// and read back as warp-broadcast LDS.128.
constexpr int LK_RT = 16;
constexpr int LK_NW = 4;
constexpr int LK_TT = LK_RT * LK_NW;  // 64 time steps per CTA tile
constexpr int LK_THREADS = LK_NW * 32;

struct LayerSmem {
  float Wd[3 * C * C];        // [(tap*64 + cin)][pos], pos gate-interleaved      48 KB
  float Wr[HALF * C];         // [k][pos]                                          8 KB
  float A[3][LK_TT][C];       //                                                  48 KB
  float G[LK_NW][LK_RT][HALF];//                                                   8 KB
  float br[C];                // gate-interleaved pairs
};

__global__ void __launch_bounds__(LK_THREADS, 2)
iaf_layer_kernel(const float* __restrict__ l_in, const float* __restrict__ cond,
                 float* __restrict__ l_out, const float* __restrict__ Wd,
                 const float* __restrict__ Wr, const float* __restrict__ br, int T, int dil,
                 int n_tiles) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LayerSmem& S = *reinterpret_cast<LayerSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < 3 * C * C / 4; i += LK_THREADS) cp_async16(&S.Wd[i * 4], Wd + i * 4, true);
  for (int i = tid; i < HALF * C / 4; i += LK_THREADS) cp_async16(&S.Wr[i * 4], Wr + i * 4, true);
  if (tid < C / 4) cp_async16(&S.br[tid * 4], br + tid * 4, true);
  cp_async_commit();

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const size_t row0 = (size_t)tile * LK_TT;
    const int t0 = (int)(row0 % T);  // tiles never straddle clips: T % LK_TT == 0
    // stage the three taps; rows before the clip start are zero (causal padding)
    for (int i = tid; i < 3 * LK_TT * (C / 4); i += LK_THREADS) {
      const int tap = i / (LK_TT * (C / 4));
      const int rem = i - tap * (LK_TT * (C / 4));
      const int r = rem >> 4, c4 = rem & 15;
      const int shift = (2 - tap) * dil;
      const bool ok = (t0 + r - shift) >= 0;
      const float* src = l_in + (ok ? (row0 + r - shift) * C + c4 * 4 : 0);
      cp_async16(&S.A[tap][r][c4 * 4], src, ok);
    }
    cp_async_commit();

    // prefetch this warp's conditioning rows (gate-interleaved float2 per lane)
    const int r0 = warp * LK_RT;
    float2 cnd[LK_RT];
#pragma unroll
    for (int i = 0; i < LK_RT; ++i)
      cnd[i] = __ldg(reinterpret_cast<const float2*>(cond + (row0 + r0 + i) * C) + lane);

    cp_async_wait_all();
    __syncthreads();

    unsigned long long acc[LK_RT];
#pragma unroll
    for (int i = 0; i < LK_RT; ++i) acc[i] = 0ull;

#pragma unroll 1
    for (int tap = 0; tap < 3; ++tap) {
#pragma unroll 2
      for (int k4 = 0; k4 < C / 4; ++k4) {
        float4 a[LK_RT];
#pragma unroll
        for (int i = 0; i < LK_RT; ++i)
          a[i] = *reinterpret_cast<const float4*>(&S.A[tap][r0 + i][k4 * 4]);
        const float* wrow = &S.Wd[(tap * C + k4 * 4) * C + 2 * lane];
        const unsigned long long w0 = lds_u64(wrow);
        const unsigned long long w1 = lds_u64(wrow + C);
        const unsigned long long w2 = lds_u64(wrow + 2 * C);
        const unsigned long long w3 = lds_u64(wrow + 3 * C);
#pragma unroll
        for (int i = 0; i < LK_RT; ++i) {
          ffma2_s(acc[i], w0, a[i].x);
          ffma2_s(acc[i], w1, a[i].y);
          ffma2_s(acc[i], w2, a[i].z);
          ffma2_s(acc[i], w3, a[i].w);
        }
      }
    }

    // gate: g = sigmoid(d[:32] + c) * tanh(d[32:] + c)   (parallel_wavenet.py:246-250)
#pragma unroll
    for (int i = 0; i < LK_RT; ++i) {
      const float2 dd = unpack2(acc[i]);
      S.G[warp][i][lane] = sigmoidf_acc(dd.x + cnd[i].x) * tanhf(dd.y + cnd[i].y);
    }
    __syncwarp();

    // residual 1x1: l_new = l + br + g @ Wr   (parallel_wavenet.py:252-254)
    const float2 b2 = *reinterpret_cast<const float2*>(&S.br[2 * lane]);
#pragma unroll
    for (int i = 0; i < LK_RT; ++i)
      acc[i] = pack2(S.A[2][r0 + i][lane] + b2.x, S.A[2][r0 + i][lane + HALF] + b2.y);
#pragma unroll 2
    for (int k4 = 0; k4 < HALF / 4; ++k4) {
      float4 g[LK_RT];
#pragma unroll
      for (int i = 0; i < LK_RT; ++i)
        g[i] = *reinterpret_cast<const float4*>(&S.G[warp][i][k4 * 4]);
      const float* wrow = &S.Wr[(k4 * 4) * C + 2 * lane];
      const unsigned long long w0 = lds_u64(wrow);
      const unsigned long long w1 = lds_u64(wrow + C);
      const unsigned long long w2 = lds_u64(wrow + 2 * C);
      const unsigned long long w3 = lds_u64(wrow + 3 * C);
#pragma unroll
      for (int i = 0; i < LK_RT; ++i) {
        ffma2_s(acc[i], w0, g[i].x);
        ffma2_s(acc[i], w1, g[i].y);
        ffma2_s(acc[i], w2, g[i].z);
        ffma2_s(acc[i], w3, g[i].w);
      }
    }
#pragma unroll
    for (int i = 0; i < LK_RT; ++i) {
      const float2 v = unpack2(acc[i]);
      float* dst = l_out + (row0 + r0 + i) * C;
      dst[lane] = v.x;
      dst[lane + HALF] = v.y;
    }
    __syncthreads();  // everyone is done with S.A before the next tile overwrites it
  }
  cp_async_wait_all();
}

// --------------------------------- head -------------------------------------
// h = relu(relu(l) @ W1 + (b1 + cond_out1));  mean = h.wm + bm;  sp = h.ws + bs
// scale = clip(softplus(sp), e^-9, e^7); log_scale = log(scale); x' = x*scale + mean
// totals: mean_tot = mean + mean_tot*scale; scale_tot *= scale; log_scale_tot += log_scale
// last flow: scale_tot = min(scale_tot, e^7); log_scale_tot = min(.,7);
//            x_out = z*scale_tot + mean_tot; optional _clip_quant_scale.
constexpr int HK_RT = 16, HK_NW = 4, HK_TT = HK_RT * HK_NW, HK_THREADS = HK_NW * 32;

struct HeadSmem {
  float W1[C * C];
  float A[HK_TT][C];
  float red[HK_NW][2][HK_RT][33];
};

struct HeadParams {
  const float* l;       // fp32 residual rows, or NULL when l_hi / l_lo (split fp16 pair) are given
  const __half* l_hi;
  const __half* l_lo;
  int cond_tiled;       // out1 plane stored row-interleaved (EpiParams::tiled_planes layout)
  const float* cond;  // out1 plane, natural channel order, biases folded
  const float* W1;
  const float* wm;
  const float* ws;
  float bm, bs;
  const float* x_in;
  const float* z;
  float* x_out;
  float* mean_tot;
  float* scale_tot;
  float* log_scale_tot;
  int first, last, quantize, use_mu_law;
  float quant_chann;
  int n_tiles;
};

__global__ void __launch_bounds__(HK_THREADS, 2) iaf_head_kernel(HeadParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  HeadSmem& S = *reinterpret_cast<HeadSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < C * C / 4; i += HK_THREADS) cp_async16(&S.W1[i * 4], p.W1 + i * 4, true);
  cp_async_commit();
  const float2 wm2 = *reinterpret_cast<const float2*>(p.wm + 2 * lane);
  const float2 ws2 = *reinterpret_cast<const float2*>(p.ws + 2 * lane);

  for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
    const size_t row0 = (size_t)tile * HK_TT;
    if (p.l) {
      for (int i = tid; i < HK_TT * (C / 4); i += HK_THREADS)
        cp_async16(&S.A[i >> 4][(i & 15) * 4], p.l + (row0 + (i >> 4)) * C + (i & 15) * 4, true);
    } else {
      // split fp16 pair (tcgen05 layer engines): l = hi + lo
      for (int i = tid; i < HK_TT * (C / 8); i += HK_THREADS) {
        const size_t off = (row0 + (i >> 3)) * C + (i & 7) * 8;
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(p.l_hi + off));
        const uint4 l = __ldg(reinterpret_cast<const uint4*>(p.l_lo + off));
        const __half2* hp = reinterpret_cast<const __half2*>(&h);
        const __half2* lp = reinterpret_cast<const __half2*>(&l);
        float o[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __half22float2(hp[e]), b = __half22float2(lp[e]);
          o[2 * e] = a.x + b.x;
          o[2 * e + 1] = a.y + b.y;
        }
        float* dst = &S.A[i >> 3][(i & 7) * 8];
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
    cp_async_commit();
    const int r0 = warp * HK_RT;
    float2 cnd[HK_RT];
    if (p.cond_tiled) {
      // element (row, col) of a row-interleaved plane: ((row/128*8 + row%128/32*2 + col/32) * 1024
      //                                                + (col%32)/4*128 + (row%32)*4 + col%4
      const size_t col_off = (size_t)(lane >> 4) * 1024 + ((lane & 15) >> 1) * 128 + (lane & 1) * 2;
#pragma unroll
      for (int i = 0; i < HK_RT; ++i) {
        const size_t grow = row0 + r0 + i;
        cnd[i] = __ldg(reinterpret_cast<const float2*>(
            p.cond + ((grow >> 7) * 8 + ((grow & 127) >> 5) * 2) * 1024 + (grow & 31) * 4 + col_off));
      }
    } else {
#pragma unroll
      for (int i = 0; i < HK_RT; ++i)
        cnd[i] = __ldg(reinterpret_cast<const float2*>(p.cond + (row0 + r0 + i) * C) + lane);
    }
    cp_async_wait_all();
    __syncthreads();

    unsigned long long acc[HK_RT];
#pragma unroll
    for (int i = 0; i < HK_RT; ++i) acc[i] = pack2(cnd[i].x, cnd[i].y);
#pragma unroll 2
    for (int k4 = 0; k4 < C / 4; ++k4) {
      float4 a[HK_RT];
#pragma unroll
      for (int i = 0; i < HK_RT; ++i) {
        a[i] = *reinterpret_cast<const float4*>(&S.A[r0 + i][k4 * 4]);
        a[i].x = fmaxf(a[i].x, 0.f); a[i].y = fmaxf(a[i].y, 0.f);
        a[i].z = fmaxf(a[i].z, 0.f); a[i].w = fmaxf(a[i].w, 0.f);
      }
      const float* wrow = &S.W1[(k4 * 4) * C + 2 * lane];
      const unsigned long long w0 = lds_u64(wrow);
      const unsigned long long w1 = lds_u64(wrow + C);
      const unsigned long long w2 = lds_u64(wrow + 2 * C);
      const unsigned long long w3 = lds_u64(wrow + 3 * C);
#pragma unroll
      for (int i = 0; i < HK_RT; ++i) {
        ffma2_s(acc[i], w0, a[i].x);
        ffma2_s(acc[i], w1, a[i].y);
        ffma2_s(acc[i], w2, a[i].z);
        ffma2_s(acc[i], w3, a[i].w);
      }
    }
#pragma unroll
    for (int i = 0; i < HK_RT; ++i) {
      const float2 h = unpack2(acc[i]);
      const float hx = fmaxf(h.x, 0.f), hy = fmaxf(h.y, 0.f);
      S.red[warp][0][i][lane] = fmaf(hx, wm2.x, hy * wm2.y);
      S.red[warp][1][i][lane] = fmaf(hx, ws2.x, hy * ws2.y);
    }
    __syncwarp();
    if (lane < HK_RT) {
      float sm = 0.f, ss = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        sm += S.red[warp][0][lane][j];
        ss += S.red[warp][1][lane][j];
      }
      const size_t row = row0 + r0 + lane;
      const float mean = sm + p.bm;
      const float sp = softplusf_acc(ss + p.bs);
      const float scale = fminf(fmaxf(sp, 1.2340980408667956e-4f /*e^-9*/), 1096.6331584284585f /*e^7*/);
      const float log_scale = logf(scale);
      const float xin = p.x_in[row];
      float mt, st, lt;
      if (p.first) {
        mt = mean; st = scale; lt = log_scale;
      } else {
        mt = fmaf(p.mean_tot[row], scale, mean);
        st = p.scale_tot[row] * scale;
        lt = p.log_scale_tot[row] + log_scale;
      }
      float xo = fmaf(xin, scale, mean);
      if (p.last) {
        st = fminf(st, 1096.6331584284585f);
        lt = fminf(lt, 7.0f);
        xo = fmaf(p.z[row], st, mt);  // new_x = x * scale_tot + mean_tot (:330)
        if (p.quantize) xo = clip_quant_scale_dev(xo, p.quant_chann, p.use_mu_law);
      }
      p.mean_tot[row] = mt;
      p.scale_tot[row] = st;
      p.log_scale_tot[row] = lt;
      p.x_out[row] = xo;
    }
    __syncthreads();
  }
  cp_async_wait_all();
}

}  // namespace

}  // namespace nsw

// =============================== host side ==================================
using namespace nsw;

struct FlowWeights {
  int L = 0;
  DevBuf start_w, start_b;  // [3][64], [64]
  DevBuf Wd, Wr, br;        // [L][192][64], [L][32][64], [L][64]  (gate-interleaved columns)
  DevBuf Wc, bc;            // cond projection B: fp32 [256][(L+1)*64]; bias [(L+1)*64]
  DevBuf Wct_hi, Wct_lo;    // fp16 [(L+1)*64][256]
  DevBuf W1, wm, ws;        // [64][64], [64], [64]
  // tcgen05 layer engine: K-major fp16 hi/lo weights + their tensor maps
  DevBuf WdT_hi, WdT_lo;    // [L][3][64 pos][64 cin]
  DevBuf WrT_hi, WrT_lo;    // [L][64 c][32 j]
  DevBuf br_nat;            // [L][64] natural channel order
  DevBuf br_cum;            // [L][64] running sum of br over layers (engine tc3 keeps the biases out of TMEM)
  alignas(64) unsigned char map_wdh[128], map_wdl[128], map_wrh[128], map_wrl[128];
  alignas(64) unsigned char map_wdh32[128], map_wdl32[128], map_wrh32[128], map_wrl32[128];  // 32-row boxes (pair kernel)
  float bm = 0.f, bs = 0.f;
  int deconv_index = 0;     // which DeconvStack feeds this flow
};

struct nsw_iaf {
  nsw_iaf_config cfg;
  int device = 0;
  int num_sms = 148;
  std::vector<FlowWeights> flows;
  // NSW_COND_ALL_FLOWS=1 (engine tc3, shared upsampling stack): the mel-cond projections of ALL flows as one GEMM (the
  // activation tile is loaded once per 128 rows instead of once per flow, one launch and one tail instead of four, no
  // column padding).  Measured 6 % fewer cycles and the same time (profiles/r02: run25, run44) for twice the plane
  // memory (1 GB at 8 x 7680), so it is off by default.
  DevBuf Wct_all_hi, Wct_all_lo, bc_all;  // the flows' Wct_* / bc back to back
  std::vector<size_t> plane_off;          // first cond plane of flow f
  size_t total_planes = 0;                // 0: one projection per flow
  std::vector<DeconvStack> deconvs;
  int max_layers = 0;
  // workspace
  int ws_B = 0, ws_F = 0;
  DevBuf mel, mel_en, mel_en_split, cond, l0, l1, x, z, mean_tot, scale_tot, log_scale_tot,
      deconv_scratch, ls0, ls1, grid_counter, sync_words, x2;
  float* x_final = nullptr;  // which of x / x2 holds the last forward's output
  // the forward of a given (B, F, quantize) replayed as a CUDA graph from its third call on (engine tc3)
  cudaGraphExec_t graph_exec = nullptr;
  int graph_B = 0, graph_F = 0, graph_quant = -1, graph_seen = 0;
  uint64_t graph_launches = 0;  // kernels inside the captured graph
  float* graph_x_final = nullptr;
  cudaEvent_t graph_ev_in = nullptr, graph_ev_out = nullptr;  // hand-over when the caller's stream is the legacy one  // ls*: fp16 hi plane then lo plane of l0 / l1 (tc2 engine)
  alignas(64) unsigned char map_act[2][2][128];  // [buffer][hi, lo]
  int map_B = 0, map_T = 0;
  // debug tap
  int tap_flow = -1, tap_layer = -1;
  float* tap_dst = nullptr;
  // profiling
  bool profiling = false;
  cudaEvent_t ev[8] = {};
  bool ev_ready = false;
  float last_ms[5] = {0, 0, 0, 0, 0};
  // pinned staging for the host entry point
  void* pin = nullptr;
  size_t pin_bytes = 0;
  cudaStream_t own_stream = nullptr;
};

static inline int gate_pos_to_channel(int pos) { return (pos & 1) ? (pos >> 1) + HALF : (pos >> 1); }

static int pack_flow(const nsw_iaf_config& cfg, const TensorMap& tm, int f, FlowWeights& fw,
                     bool want_tc) {
  const int L = cfg.num_iaf_layers[f];
  fw.L = L;
  const std::string p = "iaf_" + std::to_string(f + 1);
  const float* sw = tm.get(p + "/start_conv/W", 3 * C);
  const float* sb = tm.get(p + "/start_conv/biases", C);
  if (!sw || !sb) return NSW_EMISSING;
  NSW_TRY(upload(fw.start_w, sw, 3 * C * sizeof(float)));  // [1,3,1,64] == [3][64]
  NSW_TRY(upload(fw.start_b, sb, C * sizeof(float)));

  std::vector<float> Wd((size_t)L * 3 * C * C), Wr((size_t)L * HALF * C), br((size_t)L * C);
  const int NP = (L + 1) * C;
  std::vector<float> Wc((size_t)D * NP), bc(NP);
  for (int i = 0; i < L; ++i) {
    const std::string li = std::to_string(i + 1);
    const float* wd = tm.get(p + "/dilated_conv_" + li + "/W", 3 * C * C);  // [1,3,64,64]
    const float* bd = tm.get(p + "/dilated_conv_" + li + "/biases", C);
    const float* wc = tm.get(p + "/mel_cond_" + li + "/W", D * C);  // [1,1,256,64]
    const float* bcnd = tm.get(p + "/mel_cond_" + li + "/biases", C);
    const float* wr = tm.get(p + "/res_" + li + "/W", HALF * C);  // [1,1,32,64]
    const float* brr = tm.get(p + "/res_" + li + "/biases", C);
    if (!wd || !bd || !wc || !bcnd || !wr || !brr) return NSW_EMISSING;
    for (int k = 0; k < 3 * C; ++k)
      for (int pos = 0; pos < C; ++pos)
        Wd[((size_t)i * 3 * C + k) * C + pos] = wd[(size_t)k * C + gate_pos_to_channel(pos)];
    for (int k = 0; k < HALF; ++k)
      for (int pos = 0; pos < C; ++pos)
        Wr[((size_t)i * HALF + k) * C + pos] = wr[(size_t)k * C + gate_pos_to_channel(pos)];
    for (int pos = 0; pos < C; ++pos) br[(size_t)i * C + pos] = brr[gate_pos_to_channel(pos)];
    for (int k = 0; k < D; ++k)
      for (int pos = 0; pos < C; ++pos)
        Wc[(size_t)k * NP + i * C + pos] = wc[(size_t)k * C + gate_pos_to_channel(pos)];
    for (int pos = 0; pos < C; ++pos) {
      const int ch = gate_pos_to_channel(pos);
      bc[i * C + pos] = bd[ch] + bcnd[ch];  // dilated-conv bias folded into the cond plane
    }
  }
  const float* w1 = tm.get(p + "/out1/W", C * C);
  const float* b1 = tm.get(p + "/out1/biases", C);
  const float* wco = tm.get(p + "/mel_cond_out1/W", D * C);
  const float* bco = tm.get(p + "/mel_cond_out1/biases", C);
  const float* wm = tm.get(p + "/out2_mean/W", C);
  const float* bm = tm.get(p + "/out2_mean/biases", 1);
  const float* wsc = tm.get(p + "/out2_scale/W", C);
  const float* bsc = tm.get(p + "/out2_scale/biases", 1);
  if (!w1 || !b1 || !wco || !bco || !wm || !bm || !wsc || !bsc) return NSW_EMISSING;
  for (int k = 0; k < D; ++k)
    for (int c = 0; c < C; ++c) Wc[(size_t)k * NP + L * C + c] = wco[(size_t)k * C + c];
  for (int c = 0; c < C; ++c) bc[L * C + c] = b1[c] + bco[c];
  fw.bm = bm[0];
  fw.bs = bsc[0];
  NSW_TRY(upload(fw.Wd, Wd.data(), Wd.size() * sizeof(float)));
  NSW_TRY(upload(fw.Wr, Wr.data(), Wr.size() * sizeof(float)));
  NSW_TRY(upload(fw.br, br.data(), br.size() * sizeof(float)));
  NSW_TRY(upload(fw.Wc, Wc.data(), Wc.size() * sizeof(float)));
  NSW_TRY(upload(fw.bc, bc.data(), bc.size() * sizeof(float)));
  NSW_TRY(upload(fw.W1, w1, C * C * sizeof(float)));
  NSW_TRY(upload(fw.wm, wm, C * sizeof(float)));
  NSW_TRY(upload(fw.ws, wsc, C * sizeof(float)));
  if (cfg.engine >= NSW_ENGINE_TC2) {
    // [L][3] dilated-conv tap tiles + one more tile: W1^T of out1 (the fused head of engine tc3)
    std::vector<float> wdt(((size_t)L * 3 + 1) * C * C), wrt((size_t)L * C * HALF), brn((size_t)L * C);
    for (int i = 0; i < L; ++i) {
      const std::string li = std::to_string(i + 1);
      const float* wd = tm.get(p + "/dilated_conv_" + li + "/W", 3 * C * C);
      const float* wr = tm.get(p + "/res_" + li + "/W", HALF * C);
      const float* brr = tm.get(p + "/res_" + li + "/biases", C);
      for (int tap = 0; tap < 3; ++tap)
        for (int pos = 0; pos < C; ++pos)
          for (int k = 0; k < C; ++k)
            wdt[(((size_t)i * 3 + tap) * C + pos) * C + k] =
                wd[((size_t)tap * C + k) * C + gate_pos_to_channel(pos)];
      for (int c = 0; c < C; ++c)
        for (int j = 0; j < HALF; ++j) wrt[((size_t)i * C + c) * HALF + j] = wr[(size_t)j * C + c];
      for (int c = 0; c < C; ++c) brn[(size_t)i * C + c] = brr[c];
    }
    for (int n = 0; n < C; ++n)
      for (int k = 0; k < C; ++k) wdt[((size_t)L * 3 * C + n) * C + k] = w1[(size_t)k * C + n];
    std::vector<__half> hi(wdt.size()), lo(wdt.size());
    split_f16(wdt.data(), wdt.size(), hi.data(), lo.data());
    NSW_TRY(upload(fw.WdT_hi, hi.data(), hi.size() * 2));
    NSW_TRY(upload(fw.WdT_lo, lo.data(), lo.size() * 2));
    hi.resize(wrt.size()); lo.resize(wrt.size());
    split_f16(wrt.data(), wrt.size(), hi.data(), lo.data());
    NSW_TRY(upload(fw.WrT_hi, hi.data(), hi.size() * 2));
    NSW_TRY(upload(fw.WrT_lo, lo.data(), lo.size() * 2));
    NSW_TRY(upload(fw.br_nat, brn.data(), brn.size() * 4));
    std::vector<float> brc(brn.size());
    for (int c = 0; c < C; ++c) {
      float acc = 0.f;  // fp32 like the reference's sequential l += ... + br
      for (int i = 0; i < L; ++i) { acc += brn[(size_t)i * C + c]; brc[(size_t)i * C + c] = acc; }
    }
    NSW_TRY(upload(fw.br_cum, brc.data(), brc.size() * 4));
    NSW_TRY(layer_tc_make_weight_map(fw.map_wdh, fw.WdT_hi.as<__half>(), (L * 3 + 1) * C, C));
    NSW_TRY(layer_tc_make_weight_map(fw.map_wdl, fw.WdT_lo.as<__half>(), (L * 3 + 1) * C, C));
    NSW_TRY(layer_tc_make_weight_map(fw.map_wrh, fw.WrT_hi.as<__half>(), L * C, HALF));
    NSW_TRY(layer_tc_make_weight_map(fw.map_wrl, fw.WrT_lo.as<__half>(), L * C, HALF));
    NSW_TRY(layer_tc_make_weight_map(fw.map_wdh32, fw.WdT_hi.as<__half>(), (L * 3 + 1) * C, C, 32));
    NSW_TRY(layer_tc_make_weight_map(fw.map_wdl32, fw.WdT_lo.as<__half>(), (L * 3 + 1) * C, C, 32));
    NSW_TRY(layer_tc_make_weight_map(fw.map_wrh32, fw.WrT_hi.as<__half>(), L * C, HALF, 32));
    NSW_TRY(layer_tc_make_weight_map(fw.map_wrl32, fw.WrT_lo.as<__half>(), L * C, HALF, 32));
  }
  if (want_tc) {
    std::vector<float> bt((size_t)NP * D);
    for (int k = 0; k < D; ++k)
      for (int n = 0; n < NP; ++n) bt[(size_t)n * D + k] = Wc[(size_t)k * NP + n];
    std::vector<__half> hi(bt.size()), lo(bt.size());
    split_f16(bt.data(), bt.size(), hi.data(), lo.data());
    NSW_TRY(upload(fw.Wct_hi, hi.data(), hi.size() * sizeof(__half)));
    NSW_TRY(upload(fw.Wct_lo, lo.data(), lo.size() * sizeof(__half)));
  }
  return NSW_OK;
}

static int iaf_total_stride(const nsw_iaf_config& c) {
  int s = 1;
  for (int i = 0; i < c.num_deconv; ++i) s *= c.deconv_stride[i];
  return s;
}

extern "C" int64_t nsw_iaf_length(const nsw_iaf* h, int32_t F) {
  if (!h) return -1;
  const int64_t maxd = 1ll << (h->cfg.num_stages - 1);
  return ((int64_t)F * iaf_total_stride(h->cfg) / maxd) * maxd;
}

extern "C" int nsw_iaf_create(const nsw_iaf_config* cfg, const nsw_tensor* tensors, int32_t n,
                              int32_t device, nsw_iaf** out) {
  NSW_CHECK(cfg && tensors && out, NSW_EINVAL, "nsw_iaf_create: null argument");
  NSW_CHECK(cfg->width == C && cfg->deconv_width == D && cfg->filter_length == 3, NSW_EINVAL,
            "nsw_iaf_create: kernels are specialised for width=64, deconv_width=256, "
            "filter_length=3 (got %d, %d, %d)", cfg->width, cfg->deconv_width, cfg->filter_length);
  NSW_CHECK(cfg->num_flows >= 1 && cfg->num_flows <= NSW_MAX_FLOWS, NSW_EINVAL, "bad num_flows");
  NSW_CHECK(cfg->num_deconv >= 1 && cfg->num_deconv <= NSW_MAX_DECONV, NSW_EINVAL, "bad num_deconv");
  NSW_CHECK(cfg->num_stages >= 1 && cfg->num_stages <= 16, NSW_EINVAL, "bad num_stages");
  NSW_CHECK(cfg->loss_type == NSW_LOSS_LOGISTIC || cfg->loss_type == NSW_LOSS_GAUSS, NSW_EINVAL,
            "student loss_type must be logistic or gauss");
  NSW_CUDA(cudaSetDevice(device));
  NSW_TRY(range_guard_init(device));
  nsw_iaf* h = new nsw_iaf();
  h->cfg = *cfg;
  h->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->num_sms = prop.multiProcessorCount;
  const bool want_tc = cfg->engine >= NSW_ENGINE_TC;
  TensorMap tm(tensors, n);
  int rc = NSW_OK;
  const int n_stacks = cfg->share_deconv ? 1 : cfg->num_flows;
  h->deconvs.resize(n_stacks);
  for (int s = 0; s < n_stacks && rc == NSW_OK; ++s) {
    const std::string prefix = cfg->share_deconv ? "iaf_share/" : "iaf_" + std::to_string(s + 1) + "/";
    rc = h->deconvs[s].init(tm, prefix, cfg->num_mel, D, cfg->num_deconv, cfg->deconv_filter,
                            cfg->deconv_stride, cfg->upsample_act, want_tc);
  }
  h->flows.resize(cfg->num_flows);
  for (int f = 0; f < cfg->num_flows && rc == NSW_OK; ++f) {
    rc = pack_flow(*cfg, tm, f, h->flows[f], want_tc);
    h->flows[f].deconv_index = cfg->share_deconv ? 0 : f;
    h->max_layers = std::max(h->max_layers, cfg->num_iaf_layers[f]);
  }
  if (rc == NSW_OK && cfg->engine == NSW_ENGINE_TC3 && cfg->share_deconv && cfg->num_flows > 1 &&
      getenv("NSW_COND_ALL_FLOWS") != nullptr) {
    size_t planes = 0;
    for (int f = 0; f < cfg->num_flows; ++f) {
      h->plane_off.push_back(planes);
      planes += (size_t)h->flows[f].L + 1;
    }
    rc = h->Wct_all_hi.ensure(planes * C * D * sizeof(__half));
    if (rc == NSW_OK) rc = h->Wct_all_lo.ensure(planes * C * D * sizeof(__half));
    if (rc == NSW_OK) rc = h->bc_all.ensure(planes * C * sizeof(float));
    for (int f = 0; f < cfg->num_flows && rc == NSW_OK; ++f) {
      const FlowWeights& fw = h->flows[f];
      const size_t np = (size_t)fw.L + 1, o = h->plane_off[f];
      cudaError_t e = cudaMemcpy(h->Wct_all_hi.as<__half>() + o * C * D, fw.Wct_hi.as<__half>(), np * C * D * sizeof(__half),
                                 cudaMemcpyDeviceToDevice);
      if (e == cudaSuccess)
        e = cudaMemcpy(h->Wct_all_lo.as<__half>() + o * C * D, fw.Wct_lo.as<__half>(), np * C * D * sizeof(__half),
                       cudaMemcpyDeviceToDevice);
      if (e == cudaSuccess)
        e = cudaMemcpy(h->bc_all.as<float>() + o * C, fw.bc.as<float>(), np * C * sizeof(float), cudaMemcpyDeviceToDevice);
      if (e != cudaSuccess) {
        set_error("nsw_iaf_create: %s", cudaGetErrorString(e));
        rc = NSW_ECUDA;
      }
    }
    if (rc == NSW_OK) h->total_planes = planes;
  }
  if (rc == NSW_OK) {
    cudaError_t e = cudaFuncSetAttribute(iaf_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(LayerSmem));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(iaf_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(HeadSmem));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      set_error("nsw_iaf_create: %s", cudaGetErrorString(e));
      rc = NSW_ECUDA;
    }
  }
  if (rc != NSW_OK) {
    nsw_iaf_destroy(h);
    return rc;
  }
  *out = h;
  return NSW_OK;
}

extern "C" void nsw_iaf_destroy(nsw_iaf* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->ev_ready)
    for (auto& e : h->ev) cudaEventDestroy(e);
  if (h->pin) cudaFreeHost(h->pin);
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  if (h->graph_ev_in) cudaEventDestroy(h->graph_ev_in);
  if (h->graph_ev_out) cudaEventDestroy(h->graph_ev_out);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

static int ensure_workspace(nsw_iaf* h, int B, int F) {
  if (B <= h->ws_B && F <= h->ws_F && h->ws_B * h->ws_F >= B * F && h->ws_B > 0) return NSW_OK;
  const int nB = std::max(B, h->ws_B), nF = std::max(F, h->ws_F);
  const size_t T = (size_t)nsw_iaf_length(h, nF);
  const size_t Lc = (size_t)nF * iaf_total_stride(h->cfg);
  const size_t rows = (size_t)nB * T;
  NSW_TRY(h->mel.ensure((size_t)nB * nF * h->cfg.num_mel * sizeof(float)));
  if (h->cfg.engine >= NSW_ENGINE_TC)
    NSW_TRY(h->mel_en_split.ensure((size_t)nB * Lc * D * 2 * sizeof(__half)));
  else
    NSW_TRY(h->mel_en.ensure((size_t)nB * Lc * D * sizeof(float)));
  NSW_TRY(h->cond.ensure(std::max((size_t)h->max_layers + 1, h->total_planes) * rows * C * sizeof(float)));
  NSW_TRY(h->l0.ensure(rows * C * sizeof(float)));
  NSW_TRY(h->l1.ensure(rows * C * sizeof(float)));
  if (h->cfg.engine >= NSW_ENGINE_TC2) {
    NSW_TRY(h->sync_words.ensure((rows / 128 + 1 + (size_t)h->num_sms) * sizeof(unsigned int)));
    NSW_TRY(h->ls0.ensure(rows * C * 2 * sizeof(__half)));
    NSW_TRY(h->ls1.ensure(rows * C * 2 * sizeof(__half)));
    h->map_B = h->map_T = 0;  // buffers may have moved
  }
  NSW_TRY(h->x.ensure(rows * sizeof(float)));
  NSW_TRY(h->x2.ensure(rows * sizeof(float)));
  NSW_TRY(h->z.ensure(rows * sizeof(float)));
  NSW_TRY(h->mean_tot.ensure(rows * sizeof(float)));
  NSW_TRY(h->scale_tot.ensure(rows * sizeof(float)));
  NSW_TRY(h->log_scale_tot.ensure(rows * sizeof(float)));
  h->ws_B = nB;
  h->ws_F = nF;
  return NSW_OK;
}

extern "C" size_t nsw_iaf_workspace_bytes(const nsw_iaf* h) {
  if (!h) return 0;
  return h->mel.bytes + h->mel_en.bytes + h->mel_en_split.bytes + h->cond.bytes + h->l0.bytes +
         h->l1.bytes + h->ls0.bytes + h->ls1.bytes + h->x.bytes + h->z.bytes + h->mean_tot.bytes + h->scale_tot.bytes +
         h->log_scale_tot.bytes + h->deconv_scratch.bytes + h->x2.bytes + h->sync_words.bytes;
}

extern "C" int nsw_iaf_set_tap(nsw_iaf* h, int32_t flow, int32_t layer, float* d_l) {
  NSW_CHECK(h, NSW_EINVAL, "null handle");
  h->tap_flow = d_l ? flow : -1;
  h->tap_layer = layer;
  h->tap_dst = d_l;
  return NSW_OK;
}

extern "C" int nsw_iaf_set_profiling(nsw_iaf* h, int32_t on) {
  NSW_CHECK(h, NSW_EINVAL, "null handle");
  if (on && !h->ev_ready) {
    for (auto& e : h->ev) NSW_CUDA(cudaEventCreate(&e));
    h->ev_ready = true;
  }
  h->profiling = on != 0;
  return NSW_OK;
}

extern "C" int nsw_iaf_last_timing(nsw_iaf* h, float ms[5]) {
  NSW_CHECK(h && ms, NSW_EINVAL, "null argument");
  for (int i = 0; i < 5; ++i) ms[i] = h->last_ms[i];
  return NSW_OK;
}

extern "C" int nsw_iaf_deconv_device(nsw_iaf* h, int32_t stack, const float* d_mel, int32_t B,
                                     int32_t F, float* d_mel_en, void* stream) {
  NSW_CHECK(h && d_mel && d_mel_en, NSW_EINVAL, "null argument");
  NSW_CHECK(stack >= 0 && stack < (int)h->deconvs.size(), NSW_EINVAL, "bad deconv stack index");
  NSW_CUDA(cudaSetDevice(h->device));
  return h->deconvs[stack].forward(d_mel, B, F, d_mel_en, nullptr, nullptr, NSW_ENGINE_FFMA,
                                   h->deconv_scratch, (cudaStream_t)stream);
}

// the forward pass on internal buffers; h->mel and (optionally) h->z must already hold the inputs
static int iaf_forward_internal(nsw_iaf* h, int B, int F, bool have_z, uint64_t seed, int quantize,
                                cudaStream_t st) {
  const nsw_iaf_config& cfg = h->cfg;
  const int T = (int)nsw_iaf_length(h, F);
  NSW_CHECK(T > 0, NSW_EINVAL, "mel too short: F=%d gives length 0", F);
  NSW_CHECK(T % LK_TT == 0, NSW_EINVAL, "length %d is not a multiple of the time tile %d", T, LK_TT);
  const int Lc = F * iaf_total_stride(cfg);
  const int left = (Lc - T) / 2;  // wavenet._condition centre trim (wavenet.py:76-85)
  const size_t rows = (size_t)B * T;
  const int n_tiles = (int)(rows / LK_TT);
  const bool prof = h->profiling;
  float acc_ms[4] = {0, 0, 0, 0};
  auto rec = [&](int i) { if (prof) cudaEventRecord(h->ev[i], st); };
  auto lap = [&](int slot, int a, int b) {
    if (!prof) return;
    cudaEventSynchronize(h->ev[b]);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]);
    acc_ms[slot] += ms;
  };
  if (prof) cudaEventRecord(h->ev[6], st);

  if (!have_z) {
    const size_t n4 = (rows + 3) / 4;
    iaf_noise_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(
        h->z.as<float>(), rows, seed, cfg.loss_type == NSW_LOSS_GAUSS);
    count_launch();
  }

  const bool tc = cfg.engine >= NSW_ENGINE_TC;
  const bool tc_layers = cfg.engine >= NSW_ENGINE_TC2;
  // tc3: residual stream resident in shared memory; needs every clip to fit 4 tiles x its share of SMs
  const int tc3_clips = (cfg.engine == NSW_ENGINE_TC3 && T % 128 == 0) ? flow_tc_clips_per_launch(T, h->num_sms) : 0;
  const bool tc3 = tc3_clips > 0;
  __half* me_hi = tc ? h->mel_en_split.as<__half>() : nullptr;
  __half* me_lo = tc ? me_hi + (size_t)B * Lc * D : nullptr;
  float* l_buf[2] = {h->l0.as<float>(), h->l1.as<float>()};
  __half* ls_hi[2] = {nullptr, nullptr};
  __half* ls_lo[2] = {nullptr, nullptr};
  if (tc_layers) {
    ls_hi[0] = h->ls0.as<__half>(); ls_lo[0] = ls_hi[0] + rows * C;
    ls_hi[1] = h->ls1.as<__half>(); ls_lo[1] = ls_hi[1] + rows * C;
    if (h->map_B != B || h->map_T != T) {
      for (int b = 0; b < 2; ++b) {
        NSW_TRY(layer_tc_make_act_map(h->map_act[b][0], ls_hi[b], B, T));
        NSW_TRY(layer_tc_make_act_map(h->map_act[b][1], ls_lo[b], B, T));
      }
      h->map_B = B;
      h->map_T = T;
    }
  }
  const int grid_lk = std::min(n_tiles, 2 * h->num_sms);

  for (int f = 0; f < cfg.num_flows; ++f) {
    FlowWeights& fw = h->flows[f];
    // 1. upsampling stack (once if shared: parallel_wavenet.py:311-314, else per flow :217-220)
    if (f == 0 || !cfg.share_deconv) {
      rec(0);
      NSW_TRY(h->deconvs[fw.deconv_index].forward(h->mel.as<float>(), B, F,
                                                  tc ? nullptr : h->mel_en.as<float>(), me_hi,
                                                  me_lo, tc ? NSW_ENGINE_TC : NSW_ENGINE_FFMA, h->deconv_scratch, st));
      rec(1);
      lap(0, 0, 1);
    }
    // 2. all mel-cond projections of this flow in one GEMM, centre trim folded into a_off (tc3 with a shared
    //    upsampling stack: of all flows, before the first one)
    const bool cond_all = tc3 && h->total_planes > 0;
    float* const cond_f = h->cond.as<float>() + (cond_all ? h->plane_off[f] * rows * C : 0);
    if (cond_all) {
      if (f == 0) {
        rec(0);
        NSW_TRY(cond_proj_tc(B, Lc, T, left, (int)(h->total_planes * C), me_hi, me_lo, h->Wct_all_hi.as<__half>(),
                             h->Wct_all_lo.as<__half>(), h->bc_all.as<float>(), h->cond.as<float>(), st));
        rec(1);
        lap(1, 0, 1);
      }
    } else {
      rec(0);
      ConvGemm g;
      g.nclips = B; g.L = Lc; g.cin = D; g.ntaps = 1; g.a_off = left; g.mclip = T;
      g.N = (fw.L + 1) * C;
      EpiParams e{};
      e.mode = EPI_PLANES;
      e.bias = fw.bc.as<float>();
      e.out_f32 = h->cond.as<float>();
      if (tc3)  // activation tile resident in shared memory, every plane row-interleaved
        NSW_TRY(cond_proj_tc(B, Lc, T, left, g.N, me_hi, me_lo, fw.Wct_hi.as<__half>(), fw.Wct_lo.as<__half>(),
                             fw.bc.as<float>(), h->cond.as<float>(), st));
      else if (tc)
        NSW_TRY(conv_gemm_tc(g, me_hi, me_lo, fw.Wct_hi.as<__half>(),
                             fw.Wct_lo.as<__half>(), e, st));
      else
        NSW_TRY(conv_gemm_ffma(g, h->mel_en.as<float>(), fw.Wc.as<float>(), e, st));
      rec(1);
      lap(1, 0, 1);
    }
    // 3. start conv + residual layers
    rec(0);
    // engine tc3 fuses the start conv (reads x of rows owned by OTHER CTAs) and the head (writes x) into one
    // unsynchronised persistent kernel, so the flow's input and output must be different buffers
    float* xbuf[2] = {h->x.as<float>(), tc3 ? h->x2.as<float>() : h->x.as<float>()};
    const float* x_cur = (f == 0) ? h->z.as<float>() : xbuf[(f - 1) & 1];
    float* x_next = xbuf[f & 1];
    h->x_final = x_next;
    const bool fuse_ends = tc3 && h->tap_flow != f;  // start conv + head inside the flow kernel
    if (!fuse_ends) {
      iaf_start_conv_kernel<<<(unsigned)((rows * 16 + 255) / 256), 256, 0, st>>>(
          x_cur, l_buf[0], fw.start_w.as<float>(), fw.start_b.as<float>(), T, rows, ls_hi[0], ls_lo[0]);
      count_launch();
    }
    rec(1);
    lap(3, 0, 1);
    rec(0);
    int cur = 0;
    bool head_fused = false;
    if (h->tap_flow == f && h->tap_layer == 0)
      NSW_CUDA(cudaMemcpyAsync(h->tap_dst, l_buf[cur], rows * C * sizeof(float),
                               cudaMemcpyDeviceToDevice, st));
    if (tc_layers) {
      // one persistent cooperative launch for all layers of the flow (two when a debug tap
      // asks for the residual stream after an inner layer)
      NSW_TRY(h->grid_counter.ensure(sizeof(unsigned int)));
      const void* const maps[2][2] = {{h->map_act[0][0], h->map_act[0][1]}, {h->map_act[1][0], h->map_act[1][1]}};
      int split_at = (h->tap_flow == f && h->tap_layer >= 1 && h->tap_layer <= fw.L) ? h->tap_layer : fw.L;
      int l0 = 0;
      while (l0 < fw.L) {
        const int l1 = (l0 < split_at) ? split_at : fw.L;
        if (tc3) {
          FlowHead fh;
          fh.w_tile = fw.L * 3;
          fh.wm = fw.wm.as<float>(); fh.ws = fw.ws.as<float>(); fh.bm = fw.bm; fh.bs = fw.bs;
          fh.x_in = x_cur; fh.z = h->z.as<float>(); fh.x_out = x_next;
          fh.mean_tot = h->mean_tot.as<float>(); fh.scale_tot = h->scale_tot.as<float>();
          fh.log_scale_tot = h->log_scale_tot.as<float>();
          fh.first = (f == 0); fh.last = (f == cfg.num_flows - 1); fh.quantize = quantize;
          fh.use_mu_law = cfg.use_mu_law; fh.quant_chann = cfg.use_mu_law ? 256.0f : 65536.0f;
          head_fused = fuse_ends && l0 == 0 && l1 == fw.L;  // not when a debug tap wants this flow's rows
          FlowStart fs{x_cur, fw.start_w.as<float>(), fw.start_b.as<float>()};
          // CTA-pair variant of the flow kernel (nsw_iaf_flow_pair.cu) wherever it covers the launch: whole flow, fused
          // ends, all clips at once, an even number of tiles per clip.  NSW_FLOW_PAIR=0 keeps the single-CTA kernel
          // (read per call, so that a test can switch it).
          const char* pair_env = getenv("NSW_FLOW_PAIR");
          const bool want_pair = !(pair_env != nullptr && atoi(pair_env) == 0);
          if (want_pair && head_fused && flow_pair_pairs_per_clip(T, B, h->num_sms / 2 - 2) > 0) {
            NSW_TRY(flow_pair_launch(maps, fw.map_wdh32, fw.map_wdl32, fw.map_wrh32, fw.map_wrl32,
                                     cond_f, rows * C,
                                     fw.br_cum.as<float>(), T, B, cur, fw.L, cfg.num_stages,
                                     h->sync_words.as<unsigned int>(), h->num_sms, &fh, &fs, st));
          } else
          for (int c0 = 0; c0 < B; c0 += tc3_clips)
            NSW_TRY(flow_tc_launch(maps, fw.map_wdh, fw.map_wdl, fw.map_wrh, fw.map_wrl,
                                   cond_f + (size_t)l0 * rows * C, rows * C,
                                   fw.br_cum.as<float>(), T, c0, std::min(tc3_clips, B - c0), cur, l0, l1,
                                   cfg.num_stages, h->sync_words.as<unsigned int>(), h->num_sms,
                                   head_fused ? &fh : nullptr, head_fused ? &fs : nullptr, st));
        } else {
          NSW_TRY(layer_tc_launch(maps, fw.map_wdh, fw.map_wdl, fw.map_wrh, fw.map_wrl,
                                  cond_f + (size_t)l0 * rows * C, rows * C, ls_hi, ls_lo,
                                  fw.br_nat.as<float>(), T, (int)rows, cur, l0, l1, cfg.num_stages,
                                  h->grid_counter.as<unsigned int>(), h->num_sms, st));
        }
        cur = (cur + (l1 - l0)) & 1;
        if (l1 == split_at && split_at < fw.L + 1 && h->tap_flow == f && h->tap_layer == l1) {
          iaf_merge_split_kernel<<<(unsigned)((rows * C / 8 + 255) / 256), 256, 0, st>>>(
              ls_hi[cur], ls_lo[cur], h->tap_dst, rows * C / 8);
          count_launch();
        }
        l0 = l1;
      }
    } else {
    for (int i = 0; i < fw.L; ++i) {
      const int dil = 1 << (i % cfg.num_stages);
      iaf_layer_kernel<<<grid_lk, LK_THREADS, sizeof(LayerSmem), st>>>(
          l_buf[cur], cond_f + (size_t)i * rows * C, l_buf[cur ^ 1],
          fw.Wd.as<float>() + (size_t)i * 3 * C * C, fw.Wr.as<float>() + (size_t)i * HALF * C,
          fw.br.as<float>() + (size_t)i * C, T, dil, n_tiles);
      count_launch();
      cur ^= 1;
      if (h->tap_flow == f && h->tap_layer == i + 1)
        NSW_CUDA(cudaMemcpyAsync(h->tap_dst, l_buf[cur], rows * C * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
    }
    }
    if (tc_layers && !tc3 && fw.L > 0) {  // the head consumes fp32 rows (tc3: it reads the split pair itself)
      iaf_merge_split_kernel<<<(unsigned)((rows * C / 8 + 255) / 256), 256, 0, st>>>(
          ls_hi[cur], ls_lo[cur], l_buf[cur], rows * C / 8);
      count_launch();
    }
    rec(1);
    lap(2, 0, 1);
    // 4. head + affine + running composition
    rec(0);
    HeadParams hp;
    hp.l = (tc3 && fw.L > 0) ? nullptr : l_buf[cur];
    hp.l_hi = ls_hi[cur];
    hp.l_lo = ls_lo[cur];
    hp.cond_tiled = tc3 ? 1 : 0;
    hp.cond = cond_f + (size_t)fw.L * rows * C;
    hp.W1 = fw.W1.as<float>();
    hp.wm = fw.wm.as<float>();
    hp.ws = fw.ws.as<float>();
    hp.bm = fw.bm;
    hp.bs = fw.bs;
    hp.x_in = x_cur;
    hp.z = h->z.as<float>();
    hp.x_out = x_next;
    hp.mean_tot = h->mean_tot.as<float>();
    hp.scale_tot = h->scale_tot.as<float>();
    hp.log_scale_tot = h->log_scale_tot.as<float>();
    hp.first = (f == 0);
    hp.last = (f == cfg.num_flows - 1);
    hp.quantize = quantize;
    hp.use_mu_law = cfg.use_mu_law;
    hp.quant_chann = cfg.use_mu_law ? 256.0f : 65536.0f;
    hp.n_tiles = (int)(rows / HK_TT);
    if (!head_fused) {
      iaf_head_kernel<<<std::min(hp.n_tiles, 2 * h->num_sms), HK_THREADS, sizeof(HeadSmem), st>>>(hp);
      count_launch();
    }
    rec(1);
    lap(3, 0, 1);
  }
  NSW_CUDA(cudaGetLastError());
  if (prof) {
    cudaEventRecord(h->ev[7], st);
    cudaEventSynchronize(h->ev[7]);
    cudaEventElapsedTime(&h->last_ms[4], h->ev[6], h->ev[7]);
    for (int i = 0; i < 4; ++i) h->last_ms[i] = acc_ms[i];
  }
  return NSW_OK;
}

// Forward with launch overhead removed: the sequence of ~11 launches (memsets, cooperative kernels)
// of one (B, F, quantize) shape is captured once and replayed as a CUDA graph; noise is drawn
// outside the graph because its seed changes per call.  Anything unusual (profiling, debug tap,
// other engines, capture failure) takes the direct path.
static int iaf_forward_graphed(nsw_iaf* h, int B, int F, bool have_z, uint64_t seed, int quantize,
                               cudaStream_t st) {
  // measured gain on B200: 0.5 % of the step (the launches already queue ahead of the GPU), so it is opt-in
  static const bool use_graph = getenv("NSW_USE_GRAPH") != nullptr;
  const bool eligible = use_graph && h->cfg.engine == NSW_ENGINE_TC3 && !h->profiling && h->tap_flow < 0;
  if (!eligible) return iaf_forward_internal(h, B, F, have_z, seed, quantize, st);
  if (h->graph_B != B || h->graph_F != F || h->graph_quant != quantize) {
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    h->graph_B = B; h->graph_F = F; h->graph_quant = quantize; h->graph_seen = 0;
  }
  if (h->graph_seen >= 0) ++h->graph_seen;
  if (h->graph_seen >= 0 && h->graph_seen < 3) return iaf_forward_internal(h, B, F, have_z, seed, quantize, st);
  // the legacy default stream cannot be captured: run on the handle's own stream, ordered by events
  cudaStream_t user = st;
  const bool handover = (st == nullptr || st == cudaStreamLegacy);
  if (handover) {
    if (!h->graph_ev_in) {
      NSW_CUDA(cudaEventCreateWithFlags(&h->graph_ev_in, cudaEventDisableTiming));
      NSW_CUDA(cudaEventCreateWithFlags(&h->graph_ev_out, cudaEventDisableTiming));
    }
    st = h->own_stream;
    NSW_CUDA(cudaEventRecord(h->graph_ev_in, user));
    NSW_CUDA(cudaStreamWaitEvent(st, h->graph_ev_in, 0));
  }
  struct Rejoin {
    nsw_iaf* h; cudaStream_t user, st; bool on;
    ~Rejoin() { if (on) { cudaEventRecord(h->graph_ev_out, st); cudaStreamWaitEvent(user, h->graph_ev_out, 0); } }
  } rejoin{h, user, st, handover};
  if (!have_z) {
    const size_t rows = (size_t)B * nsw_iaf_length(h, F);
    const size_t n4 = (rows + 3) / 4;
    iaf_noise_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(h->z.as<float>(), rows, seed,
                                                                    h->cfg.loss_type == NSW_LOSS_GAUSS);
    count_launch();
  }
  if (!h->graph_exec && h->graph_seen >= 0) {
    cudaGraph_t graph = nullptr;
    const uint64_t l0 = g_launch_count.load();
    bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    int rc = NSW_OK;
    if (ok) {
      rc = iaf_forward_internal(h, B, F, true, seed, quantize, st);
      ok = cudaStreamEndCapture(st, &graph) == cudaSuccess && rc == NSW_OK && graph != nullptr;
    }
    h->graph_launches = g_launch_count.load() - l0;
    g_launch_count.fetch_sub(h->graph_launches);  // counted when the graph is launched
    if (ok) ok = cudaGraphInstantiate(&h->graph_exec, graph, 0) == cudaSuccess;
    if (graph) cudaGraphDestroy(graph);
    if (!ok) {
      cudaGetLastError();
      h->graph_exec = nullptr;
      h->graph_seen = -1;  // do not try again for this shape
      return iaf_forward_internal(h, B, F, true, seed, quantize, st);
    }
    h->graph_x_final = h->x_final;
  }
  if (h->graph_exec) {
    NSW_CUDA(cudaGraphLaunch(h->graph_exec, st));
    count_launch((int)h->graph_launches);
    h->x_final = h->graph_x_final;
    return NSW_OK;
  }
  return iaf_forward_internal(h, B, F, true, seed, quantize, st);
}

extern "C" int nsw_iaf_forward_device(nsw_iaf* h, const float* d_mel, const float* d_z,
                                      uint64_t seed, int32_t B, int32_t F, int32_t quantize,
                                      float* d_x, float* d_mean_tot, float* d_scale_tot,
                                      float* d_log_scale_tot, float* d_rand_input, void* stream) {
  NSW_CHECK(h && d_mel, NSW_EINVAL, "null argument");
  NSW_CHECK(B >= 1 && F >= 1, NSW_EINVAL, "bad batch/frames %d/%d", B, F);
  NSW_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  NSW_TRY(ensure_workspace(h, B, F));
  const size_t rows = (size_t)B * nsw_iaf_length(h, F);
  NSW_CUDA(cudaMemcpyAsync(h->mel.p, d_mel, (size_t)B * F * h->cfg.num_mel * sizeof(float),
                           cudaMemcpyDeviceToDevice, st));
  if (d_z) NSW_CUDA(cudaMemcpyAsync(h->z.p, d_z, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
  NSW_TRY(iaf_forward_graphed(h, B, F, d_z != nullptr, seed, quantize, st));
  const size_t nb = rows * sizeof(float);
  if (d_x) NSW_CUDA(cudaMemcpyAsync(d_x, h->x_final, nb, cudaMemcpyDeviceToDevice, st));
  if (d_mean_tot) NSW_CUDA(cudaMemcpyAsync(d_mean_tot, h->mean_tot.p, nb, cudaMemcpyDeviceToDevice, st));
  if (d_scale_tot) NSW_CUDA(cudaMemcpyAsync(d_scale_tot, h->scale_tot.p, nb, cudaMemcpyDeviceToDevice, st));
  if (d_log_scale_tot)
    NSW_CUDA(cudaMemcpyAsync(d_log_scale_tot, h->log_scale_tot.p, nb, cudaMemcpyDeviceToDevice, st));
  if (d_rand_input) NSW_CUDA(cudaMemcpyAsync(d_rand_input, h->z.p, nb, cudaMemcpyDeviceToDevice, st));
  return NSW_OK;
}

extern "C" int nsw_iaf_forward_host(nsw_iaf* h, const float* mel, const float* z, uint64_t seed,
                                    int32_t B, int32_t F, int32_t quantize, float* x,
                                    float* mean_tot, float* scale_tot, float* log_scale_tot,
                                    float* rand_input) {
  NSW_CHECK(h && mel, NSW_EINVAL, "null argument");
  NSW_CHECK(B >= 1 && F >= 1, NSW_EINVAL, "bad batch/frames %d/%d", B, F);
  NSW_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = h->own_stream;
  NSW_TRY(ensure_workspace(h, B, F));
  const size_t rows = (size_t)B * nsw_iaf_length(h, F);
  NSW_CUDA(cudaMemcpyAsync(h->mel.p, mel, (size_t)B * F * h->cfg.num_mel * sizeof(float),
                           cudaMemcpyHostToDevice, st));
  if (z) NSW_CUDA(cudaMemcpyAsync(h->z.p, z, rows * sizeof(float), cudaMemcpyHostToDevice, st));
  NSW_TRY(iaf_forward_graphed(h, B, F, z != nullptr, seed, quantize, st));
  const size_t nb = rows * sizeof(float);
  if (x) NSW_CUDA(cudaMemcpyAsync(x, h->x_final, nb, cudaMemcpyDeviceToHost, st));
  if (mean_tot) NSW_CUDA(cudaMemcpyAsync(mean_tot, h->mean_tot.p, nb, cudaMemcpyDeviceToHost, st));
  if (scale_tot) NSW_CUDA(cudaMemcpyAsync(scale_tot, h->scale_tot.p, nb, cudaMemcpyDeviceToHost, st));
  if (log_scale_tot)
    NSW_CUDA(cudaMemcpyAsync(log_scale_tot, h->log_scale_tot.p, nb, cudaMemcpyDeviceToHost, st));
  if (rand_input) NSW_CUDA(cudaMemcpyAsync(rand_input, h->z.p, nb, cudaMemcpyDeviceToHost, st));
  NSW_CUDA(cudaStreamSynchronize(st));
  return range_check("nsw_iaf_forward_host");
}

NSW_RANGE_GUARD_TU(iaf)
