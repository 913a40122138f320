// fp32 CUDA-core engine of the conv-GEMM (see nsw_gemm.cuh) and the transposed-conv
// upsampling stack built on it.  This is the bit-faithful fp32 path; the tcgen05
// engine in nsw_gemm_tc.cu is the fast one.
#include "nsw_gemm.cuh"

namespace nsw {

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ void store_epi(const ConvGemm& g, const EpiParams& e, int r, int n,
                                          float4 v) {
  const int M = g.nclips * g.mclip;
  if (r >= M) return;
  if (e.mode == EPI_PLANES) {
    const float4 b = *reinterpret_cast<const float4*>(e.bias + n);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    const size_t plane = (size_t)(n >> 6);
    float* dst = e.out_f32 + (plane * (size_t)M + (size_t)r) * 64 + (n & 63);
    *reinterpret_cast<float4*>(dst) = v;
  } else if (e.mode == EPI_ROWS) {
    // plain row-major result (+ bias): out[row * ld_out + n]
    const float4 b = *reinterpret_cast<const float4*>(e.bias + n);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    *reinterpret_cast<float4*>(e.out_f32 + (size_t)r * e.ld_out + n) = v;
  } else {
    const int rpc = e.flat_rows > 0 ? e.flat_rows : g.mclip;  // flattened input: the GEMM's "clip" is all clips
    const int clip = r / rpc, m = r - clip * rpc;
    const int rr = n / e.cout, co = n - rr * e.cout;
    const int o = m * e.s + rr - e.p;
    if (o < 0 || o >= e.Lout) return;
    const float4 b = *reinterpret_cast<const float4*>(e.bias + co);
    v.x = apply_act(v.x + b.x, e.act);
    v.y = apply_act(v.y + b.y, e.act);
    v.z = apply_act(v.z + b.z, e.act);
    v.w = apply_act(v.w + b.w, e.act);
    const size_t off = ((size_t)clip * (e.out_clip_rows > 0 ? e.out_clip_rows : e.Lout) + o) * e.cout + co;
    if (e.out_f32) *reinterpret_cast<float4*>(e.out_f32 + off) = v;
    if (e.out_hi) {
      float f[4] = {v.x, v.y, v.z, v.w};
      __half hi[4], lo[4];
      uint32_t rmx = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        range_track(rmx, f[i]);
        hi[i] = __float2half_rn(f[i]);
        lo[i] = __float2half_rn(f[i] - __half2float(hi[i]));
      }
      range_commit(rmx);
      *reinterpret_cast<uint2*>(e.out_hi + off) = *reinterpret_cast<uint2*>(hi);
      *reinterpret_cast<uint2*>(e.out_lo + off) = *reinterpret_cast<uint2*>(lo);
    }
  }
}

__global__ void __launch_bounds__(256)
conv_gemm_ffma_kernel(ConvGemm g, const float* __restrict__ X, const float* __restrict__ Bw,
                      EpiParams e) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int M = g.nclips * g.mclip, K = g.ntaps * g.cin;

  const int a_r = tid >> 2, a_kq = tid & 3;
  const int arow = m0 + a_r;
  const bool arow_ok = arow < M;
  const int aclip = arow_ok ? arow / g.mclip : 0;
  const int am = arow - aclip * g.mclip;
  const int b_k = tid >> 4, b_n4 = tid & 15;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    const int k = k0 + a_kq * 4;
    const int tap = k / g.cin, c = k - tap * g.cin;
    const int frame = am + g.a_off + tap;
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (arow_ok && frame >= 0 && frame < g.L)
      av = *reinterpret_cast<const float4*>(X + ((size_t)aclip * g.L + frame) * g.cin + c);
    As[a_kq * 4 + 0][a_r] = av.x;
    As[a_kq * 4 + 1][a_r] = av.y;
    As[a_kq * 4 + 2][a_r] = av.z;
    As[a_kq * 4 + 3][a_r] = av.w;
    *reinterpret_cast<float4*>(&Bs[b_k][b_n4 * 4]) =
        *reinterpret_cast<const float4*>(Bw + (size_t)(k0 + b_k) * g.N + n0 + b_n4 * 4);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float af[4] = {a.x, a.y, a.z, a.w};
      const float bf[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    store_epi(g, e, m0 + ty * 4 + i, n0 + tx * 4,
              make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
}

}  // namespace

int conv_gemm_ffma(const ConvGemm& g, const float* X, const float* Bw, const EpiParams& e,
                   cudaStream_t stream) {
  const int M = g.nclips * g.mclip, K = g.ntaps * g.cin;
  NSW_CHECK(g.N % BN == 0 && K % BK == 0 && g.cin % 4 == 0, NSW_EINVAL,
            "conv_gemm_ffma: unsupported shape N=%d K=%d cin=%d", g.N, K, g.cin);
  dim3 grid(g.N / BN, (M + BM - 1) / BM);
  conv_gemm_ffma_kernel<<<grid, 256, 0, stream>>>(g, X, Bw, e);
  count_launch();
  NSW_CUDA(cudaGetLastError());
  return NSW_OK;
}

void split_f16(const float* src, size_t n, __half* hi, __half* lo) {
  for (size_t i = 0; i < n; ++i) {
    hi[i] = __float2half_rn(src[i]);
    lo[i] = __float2half_rn(src[i] - __half2float(hi[i]));
  }
}

// mel [B, F, cin] fp32 -> split fp16 [B, rows, cp] with zero channels cin..cp-1 and zero rows F..rows-1
namespace {
__global__ void deconv_pad_split_kernel(const float* __restrict__ x, __half* __restrict__ hi, __half* __restrict__ lo, int B,
                                        int F, int cin, int rows, int cp) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n = (size_t)B * rows * cp;
  if (i >= n) return;
  const int c = (int)(i % cp);
  const size_t br = i / cp;
  const int r = (int)(br % rows), b = (int)(br / rows);
  const float v = (c < cin && r < F) ? x[((size_t)b * F + r) * cin + c] : 0.f;
  const __half h = __float2half_rn(v);
  hi[i] = h;
  lo[i] = __float2half_rn(v - __half2float(h));
}
}  // namespace

// ------------------------------- DeconvStack -------------------------------
int DeconvStack::init(const TensorMap& tm, const std::string& prefix, int num_mel, int width,
                      int n, const int32_t* filt, const int32_t* stride, int act_kind,
                      bool want_tc) {
  act = act_kind;
  total_stride = 1;
  layers.clear();
  layers.resize(n);
  int cin = num_mel;
  for (int i = 0; i < n; ++i) {
    DeconvLayer& L = layers[i];
    DeconvGeom& g = L.g;
    g.k = filt[i];
    g.s = stride[i];
    g.cin = cin;
    g.cout = width;
    NSW_CHECK(g.s > 0 && g.k >= 1, NSW_EINVAL, "deconv layer %d: bad filter %d / stride %d", i + 1, g.k, g.s);
    total_stride *= g.s;
    // which upsampler the checkpoint holds decides: resize_conv_i/{W,biases} (use_resize_conv, wavenet.py:37-39) or
    // trans_conv_i/{kernel,bias}
    const std::string rbase = prefix + "resize_conv_" + std::to_string(i + 1);
    const bool resize = tm.has(rbase + "/W");
    const float* bias = nullptr;
    std::vector<float> bw;
    if (resize) {
      // masked.resize_conv1d (masked.py:294-322): nearest-neighbour upsampling by s (x_up[u] = x[u / s]), then a
      // non-causal SAME conv of length k: y[o] = b + sum_j x_up[o + j - pl] W[j], pl = (k - 1) / 2.  For output phase
      // r = o mod s the taps that fall on input frame m + q add up to ONE effective kernel, so the layer is the same
      // conv-GEMM as the transposed conv: frames m + qmin .. m + qmax against phase-folded weights.
      const int pl = (g.k - 1) / 2;
      auto fdiv = [](int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); };
      const int qmin = fdiv(-pl, g.s), qmax = fdiv(g.s - 1 + g.k - 1 - pl, g.s);
      g.ntaps = qmax - qmin + 1;
      g.a_off = qmin;
      g.p = 0;
      const float* w = tm.get(rbase + "/W", (int64_t)g.k * g.cin * g.cout);  // [1,k,cin,cout]
      bias = tm.get(rbase + "/biases", g.cout);
      if (!w || !bias) return NSW_EMISSING;
      const int N = g.s * g.cout;
      std::vector<double> acc((size_t)g.ntaps * g.cin * N, 0.0);  // up to s taps fold into one weight: sum in fp64
      for (int r = 0; r < g.s; ++r)
        for (int j = 0; j < g.k; ++j) {
          const int tp = fdiv(r + j - pl, g.s) - qmin;
          for (int c = 0; c < g.cin; ++c)
            for (int co = 0; co < g.cout; ++co)
              acc[(size_t)(tp * g.cin + c) * N + r * g.cout + co] += (double)w[((size_t)j * g.cin + c) * g.cout + co];
        }
      bw.resize(acc.size());
      for (size_t q = 0; q < acc.size(); ++q) bw[q] = (float)acc[q];
    } else {
      NSW_CHECK(g.k % g.s == 0 && g.k >= g.s, NSW_EINVAL,
                "deconv layer %d: filter %d must be a multiple of stride %d", i + 1, g.k, g.s);
      g.ntaps = g.k / g.s;
      g.a_off = -(g.ntaps - 1);
      g.p = (g.k - g.s) / 2;  // TF conv2d_transpose SAME: left pad of the forward conv
      const std::string base = prefix + "trans_conv_" + std::to_string(i + 1);
      const float* kern = tm.get(base + "/kernel", (int64_t)g.k * g.cout * g.cin);  // [1,k,cout,cin]
      bias = tm.get(base + "/bias", g.cout);
      if (!kern || !bias) return NSW_EMISSING;
      const int N = g.s * g.cout;
      bw.resize((size_t)g.ntaps * g.cin * N);
      for (int tp = 0; tp < g.ntaps; ++tp) {
        const int q = g.ntaps - 1 - tp;  // X frame m-q  <->  A tap index tp (a_off = -(ntaps-1))
        for (int c = 0; c < g.cin; ++c)
          for (int r = 0; r < g.s; ++r) {
            const int j = r + q * g.s;
            for (int co = 0; co < g.cout; ++co)
              bw[(size_t)(tp * g.cin + c) * N + r * g.cout + co] =
                  kern[((size_t)j * g.cout + co) * g.cin + c];
          }
      }
    }
    NSW_CHECK(g.cin % 4 == 0 && g.cout % 64 == 0 && (g.ntaps * g.cin) % 16 == 0, NSW_EINVAL,
              "deconv layer %d: unsupported channel counts %d -> %d", i + 1, g.cin, g.cout);
    const int K = g.ntaps * g.cin, N = g.s * g.cout;
    NSW_TRY(upload(L.Bw, bw.data(), bw.size() * sizeof(float)));
    NSW_TRY(upload(L.bias, bias, g.cout * sizeof(float)));
    L.cin_pad = (g.cin + 63) / 64 * 64;
    if (want_tc) {
      // K-major, channels padded to a multiple of 64 with zero columns (layer 1: 80 mel bins -> 128), so that every
      // layer of the stack can run on the tensor cores
      const int cp = L.cin_pad, Kp = g.ntaps * cp;
      std::vector<float> bt((size_t)N * Kp, 0.f);
      for (int tp = 0; tp < g.ntaps; ++tp)
        for (int c = 0; c < g.cin; ++c)
          for (int nn = 0; nn < N; ++nn) bt[(size_t)nn * Kp + tp * cp + c] = bw[(size_t)(tp * g.cin + c) * N + nn];
      std::vector<__half> hi(bt.size()), lo(bt.size());
      split_f16(bt.data(), bt.size(), hi.data(), lo.data());
      NSW_TRY(upload(L.Bt_hi, hi.data(), hi.size() * sizeof(__half)));
      NSW_TRY(upload(L.Bt_lo, lo.data(), lo.size() * sizeof(__half)));
    }
    cin = width;
  }
  return NSW_OK;
}

int DeconvStack::forward(const float* d_mel, int B, int F, float* out_f32, __half* out_hi,
                         __half* out_lo, int engine, DevBuf& scratch,
                         cudaStream_t stream) const {
  const int n = (int)layers.size();
  NSW_CHECK(n >= 1, NSW_EINVAL, "deconv stack is empty");
  // intermediate activations: layer i output has length Li = F * prod(s_0..s_i)
  // zero rows the NEXT layer needs after each clip of layer i's output to run over all clips as one flattened clip
  auto pad_rows = [&](int i, int Lout) {
    const DeconvGeom& nx = layers[i + 1].g;
    return std::max(0, std::max(-nx.a_off, nx.mclip(Lout) + nx.a_off + nx.ntaps - 1 - Lout));
  };
  size_t max_elems = 0;
  {
    int L = F;
    for (int i = 0; i + 1 < n; ++i) {
      L *= layers[i].g.s;
      max_elems = std::max(max_elems, (size_t)B * (L + pad_rows(i, L)) * layers[i].g.cout);
    }
  }
  // tensor-core engine: the mel input itself as a padded split pair, so that layer 1 runs on tensor cores too
  const DeconvGeom& g0 = layers[0].g;
  const int cp0 = layers[0].cin_pad;
  const int P0 = std::max(0, std::max(-g0.a_off, g0.mclip(F) + g0.a_off + g0.ntaps - 1 - F));
  const bool in_tc = engine == NSW_ENGINE_TC && layers[0].Bt_hi.p != nullptr && (g0.s * g0.cout) % 64 == 0;
  const size_t in_elems = in_tc ? (size_t)B * (F + P0) * cp0 : 0;  // halves per plane
  // two ping-pong slots, each big enough for fp32 or (fp16 hi + fp16 lo), + the split input
  NSW_TRY(scratch.ensure(2 * max_elems * sizeof(float) + 256 + 2 * in_elems * sizeof(__half)));
  float* slot[2] = {scratch.as<float>(), scratch.as<float>() + max_elems};

  const float* x_f32 = d_mel;
  const __half *x_hi = nullptr, *x_lo = nullptr;
  int L = F;
  int in_rows = 0;  // > 0: the input of this layer is padded to in_rows rows per clip (zero rows after the data)
  if (in_tc) {
    __half* ih = reinterpret_cast<__half*>(scratch.as<float>() + 2 * max_elems + 64);
    __half* il = ih + in_elems;
    deconv_pad_split_kernel<<<(unsigned)((in_elems + 255) / 256), 256, 0, stream>>>(d_mel, ih, il, B, F, g0.cin, F + P0, cp0);
    count_launch();
    x_hi = ih;
    x_lo = il;
    in_rows = F + P0;
  }
  for (int i = 0; i < n; ++i) {
    const DeconvLayer& ly = layers[i];
    const DeconvGeom& dg = ly.g;
    const bool last = (i == n - 1);
    ConvGemm g;
    g.nclips = B;
    g.L = L;
    g.cin = (engine == NSW_ENGINE_TC && x_hi != nullptr) ? ly.cin_pad : dg.cin;
    g.ntaps = dg.ntaps;
    g.a_off = dg.a_off;
    g.mclip = dg.mclip(L);
    g.N = dg.s * dg.cout;
    EpiParams e{};
    e.mode = EPI_DECONV;
    e.bias = ly.bias.as<float>();
    e.s = dg.s;
    e.p = dg.p;
    e.cout = dg.cout;
    e.act = act;
    e.Lout = L * dg.s;
    const bool this_tc = (engine == NSW_ENGINE_TC) && x_hi != nullptr && conv_gemm_tc_supported(g);
    // does the NEXT consumer want split-fp16 input?
    bool next_wants_split = false;
    if (engine == NSW_ENGINE_TC) {
      if (last) {
        next_wants_split = (out_hi != nullptr);
      } else {
        ConvGemm gn = g;
        gn.L = e.Lout;
        gn.cin = layers[i + 1].g.cin;
        gn.ntaps = layers[i + 1].g.ntaps;
        gn.N = layers[i + 1].g.s * layers[i + 1].g.cout;
        next_wants_split = conv_gemm_tc_supported(gn);
      }
    }
    float* o_f32 = nullptr;
    __half *o_hi = nullptr, *o_lo = nullptr;
    if (last) {
      o_f32 = out_f32;
      if (next_wants_split) { o_hi = out_hi; o_lo = out_lo; }
    } else {
      float* s = slot[i & 1];
      if (next_wants_split) {
        o_hi = reinterpret_cast<__half*>(s);
        o_lo = o_hi + (size_t)B * e.Lout * dg.cout;
      } else {
        o_f32 = s;
      }
    }
    NSW_CHECK(o_f32 || o_hi, NSW_EINVAL, "deconv: no output buffer for layer %d", i + 1);
    // an intermediate that the next layer reads as split fp16 on tensor cores is written padded (see EpiParams)
    int out_rows = 0;
    if (!last && next_wants_split && !o_f32) {
      const int P = pad_rows(i, e.Lout);
      out_rows = e.Lout + P;
      o_lo = o_hi + (size_t)B * out_rows * dg.cout;
      e.out_clip_rows = out_rows;
      if (P > 0) {
        const size_t pitch = (size_t)out_rows * dg.cout * sizeof(__half), width = (size_t)P * dg.cout * sizeof(__half);
        NSW_CUDA(cudaMemset2DAsync(o_hi + (size_t)e.Lout * dg.cout, pitch, 0, width, B, stream));
        NSW_CUDA(cudaMemset2DAsync(o_lo + (size_t)e.Lout * dg.cout, pitch, 0, width, B, stream));
      }
    }
    e.out_f32 = o_f32;
    e.out_hi = o_hi;
    e.out_lo = o_lo;
    if (this_tc) {
      if (in_rows > 0) {  // all clips as one flattened clip
        g.nclips = 1;
        g.L = B * in_rows;
        g.mclip = B * in_rows;
        e.flat_rows = in_rows;
      }
      NSW_TRY(conv_gemm_tc(g, x_hi, x_lo, ly.Bt_hi.as<__half>(),
                           ly.Bt_lo.as<__half>(), e, stream));
    } else {
      NSW_CHECK(x_f32 != nullptr, NSW_EINVAL, "deconv layer %d: fp32 input unavailable", i + 1);
      NSW_TRY(conv_gemm_ffma(g, x_f32, ly.Bw.as<float>(), e, stream));
    }
    x_f32 = o_f32;
    x_hi = o_hi;
    x_lo = o_lo;
    L = e.Lout;
    in_rows = out_rows;
  }
  return NSW_OK;
}

}  // namespace nsw

NSW_RANGE_GUARD_TU(gemm_ffma)
