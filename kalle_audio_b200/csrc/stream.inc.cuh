// Stateful streaming decoder (included by kvae.cu inside its anonymous-namespace context).
//
// The reference streams with decode_audio(chunked=True, chunk_size=128, overlap=32)
// (stable_audio_tools/models/autoencoders.py:499-560): overlapping windows are decoded from scratch and the middle
// of each is pasted into the output -- 1.33x the work, and every window pays the zero-padded edges.  Here every layer
// keeps the tail of its own input between calls instead (persistent per-layer halo state), so each output row of
// each layer is computed exactly once: 0 % recompute, and the stream equals the unchunked decode bit for bit.
//
// Mechanics.  Every convolution is run in its "valid" form on a window buffer [carry rows | new rows]:
//   * a layer whose taps read input rows q + d, d in [dmin, dmax], keeps span = dmax - dmin rows between calls and
//     produces out = rows - span new outputs per call, trailing its input by dmax rows (that accumulated lag is the
//     decoder's look-ahead, 10 latent frames for the SAO / 12.5 Hz strides);
//   * the kernels are the ones of the batch path: the window is just another tensor map, all tap offsets are shifted
//     by -dmin so they are non-negative (StreamGeom::row_bias), outputs are appended to the consumer's window buffer;
//   * at the start of a stream a buffer holds -dmin rows of zeros (the conv's left padding); at the end
//     (kvae_decode_stream_end) every layer emits dmax more rows, reading past its last row, where TMA's out-of-bounds
//     zero fill is the conv's right padding;
//   * after a call the last `span` rows of every buffer move to its front (one batched copy kernel).
// A call with the same buffer fill and frame count as an earlier one re-uses its prepared launches, and (optionally)
// replays them as one CUDA graph: the steady state of a constant-hop stream is one graph launch per hop.

struct StreamOut {            // geometry shared by the raw (stream) and act (operand) tensors one step produces
  long long rows = 0;         // valid rows per clip currently in the buffers
  long long abs_base = 0;     // absolute index (at this tensor's rate) of buffer row 0
  long long cap = 0;          // rows per clip the buffers can hold (= pitch)
  int keep = 0;               // rows the consumer needs between calls (its span)
  int init_rows = 0;          // zero rows at the start of a stream (the consumer's -dmin)
  size_t off_raw = 0, off_act = 0;
  int raw_row_bytes = 0, act_row_bytes = 0;   // 0: tensor not produced
};

struct StreamStep {
  int kind = -1;              // 0 tensor-core conv, 4 fused ResidualUnit (this step + next), 5 second half of a fused unit,
                              // 6 tensor-core tail, 3 CUDA-core tail
  int bias = 0, span = 0, dmax = 0;   // in input rows
  int P_out = 1;
  int in_out = -1;            // index of the StreamOut this step reads (-1: the stream's input buffer, slot 0 of `outs`)
};

struct StreamCarry { uint8_t* dst; const uint8_t* src; unsigned long long bytes; unsigned long long pitch; };

__global__ void stream_carry_kernel(const StreamCarry* d, int n_desc) {
  const StreamCarry c = d[blockIdx.y];
  uint8_t* dst = c.dst + static_cast<size_t>(blockIdx.z) * c.pitch;
  const uint8_t* src = c.src + static_cast<size_t>(blockIdx.z) * c.pitch;
  const size_t n16 = c.bytes / 16;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n16; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  (void)n_desc;
}

struct PreparedPush {
  std::vector<int> launch;                 // per step: 1 = launch
  std::vector<ConvLaunch2> umma2;
  std::vector<RuLaunch> ru;
  PreparedRun::WaveOutTc wo_tc;
  WaveOutParams wo_cc;
  int wo_cout = 0;
  long long n_samples = 0;
  int in_rows0 = 0;                        // row of the input buffer where the new frames go
  std::vector<StreamCarry> carry;          // non-overlapping moves (one kernel)
  std::vector<StreamCarry> carry_overlap;  // overlapping moves (through scratch)
  StreamCarry* carry_dev = nullptr;
  std::vector<StreamOut> after;            // state after this call
  cudaGraphExec_t graph = nullptr;
  bool eager_done = false;
  ~PreparedPush() {
    if (graph) cudaGraphExecDestroy(graph);
    if (carry_dev) cudaFree(carry_dev);
  }
};

}  // namespace

struct kvae_stream {
  kvae_plan* plan = nullptr;
  int B = 0, max_frames = 0;
  bool use_graphs = false;
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  std::vector<StreamOut> outs;      // [0] = channels-last copy of the latent input, [1 + k] = outputs of step k
  std::vector<StreamStep> ss;
  uint8_t* z_stage = nullptr;       // [B, latent, max_frames] in the caller's dtype (4 bytes per element reserved)
  float* wav_stage = nullptr;       // [B, io, max_samples]
  uint8_t* scratch = nullptr;       // overlapping carry moves
  size_t scratch_bytes = 0;
  long long max_samples = 0;
  long long ratio = 1;
  bool stream_f16 = false;
  std::map<std::vector<long long>, std::unique_ptr<PreparedPush>> cache;
  long long frames_in = 0, samples_out = 0;
  bool needs_zero = false;          // the carry regions must be zeroed before the next stream starts
  cudaStream_t cap_stream = nullptr;   // graph capture happens here (the caller's stream may be the legacy default)
};

namespace {

void stream_reset(kvae_stream* s) {
  for (StreamOut& o : s->outs) {
    o.rows = o.init_rows;
    o.abs_base = -static_cast<long long>(o.init_rows);
  }
  s->frames_in = 0;
  s->samples_out = 0;
}

// per-step tap statistics of the decoder's convs (all stride-1 convs or transposed convs)
bool stream_step_geometry(const ConvGeom& g, StreamStep& st, std::string& err) {
  TapPlan tp;
  if (!build_taps(g, false, tp, err)) return false;
  if (tp.P_in != 1) { err = "streaming: strided convolutions (encoders) are not supported"; return false; }
  int dmin = 1 << 30, dmax = -(1 << 30);
  for (const Tap& t : tp.taps) {
    dmin = std::min(dmin, t.a_row + t.shift);
    dmax = std::max(dmax, t.a_row + t.shift);
  }
  st.bias = -dmin;
  st.dmax = dmax;
  st.span = dmax - dmin;
  st.P_out = tp.P_out;
  if (dmin > 0 || dmax < 0) { err = "streaming: one-sided receptive field unsupported"; return false; }
  return true;
}

// Lays out one call: how many rows every step produces, the launches, the carry moves and the state afterwards.
bool stream_prepare(kvae_stream* s, int n_frames, bool flush, int z_dtype, PreparedPush& P, std::string& err) {
  kvae_plan* p = s->plan;
  const std::vector<Step>& steps = p->steps;
  const int n = static_cast<int>(steps.size());
  const int B = s->B;
  const int sf16 = s->stream_f16 ? 1 : 0;
  const int split = (p->precision == KVAE_PREC_F32) ? 2 : 1;
  std::vector<StreamOut> cur = s->outs;
  P.launch.assign(n, 0);
  P.umma2.resize(n);
  P.ru.resize(n);
  P.in_rows0 = static_cast<int>(cur[0].rows);
  cur[0].rows += n_frames;
  if (cur[0].rows > cur[0].cap) { err = "streaming: input buffer overflow"; return false; }
  uint8_t* base = s->ws;
  auto raw_ptr = [&](const StreamOut& o, long long row) -> uint8_t* { return base + o.off_raw + static_cast<size_t>(row) * o.raw_row_bytes; };
  auto act_ptr = [&](const StreamOut& o, long long row) -> uint8_t* { return base + o.off_act + static_cast<size_t>(row) * o.act_row_bytes; };
  P.n_samples = 0;
  for (int k = 0; k < n; ++k) {
    const StreamStep& st = s->ss[k];
    if (st.kind == 5) continue;
    const Step& sp = steps[k];
    const ConvLayer& c = p->convs[sp.conv];
    StreamOut& in = cur[st.in_out + 1];
    const int ko = (st.kind == 4) ? k + 1 : k;              // the step whose tensors receive the result
    StreamOut& out = cur[ko + 1];
    const bool last = (ko == n - 1);
    long long out_q = in.rows - st.span + (flush ? st.dmax : 0);
    if (out_q <= 0) continue;
    const long long out_rows = out_q * st.P_out;
    const long long out_abs = (in.abs_base + st.bias) * st.P_out;
    if (!last) {
      if (out.abs_base + out.rows != out_abs) { err = "internal: streaming row bookkeeping at step " + std::to_string(k); return false; }
      if (out.rows + out_rows > out.cap) { err = "streaming: window buffer overflow at step " + std::to_string(k); return false; }
    }
    StreamGeom sg;
    sg.in_rows = static_cast<int>(in.rows);
    sg.out_q = static_cast<int>(out_q);
    sg.row_bias = st.bias;
    sg.in_pitch = in.cap;
    sg.raw_pitch = out.cap;
    sg.act_pitch = out.cap;
    const Step& so = steps[ko];
    // skip connection: rows of the residual-source stream aligned with this step's output rows
    const uint8_t* res = nullptr;
    if (so.residual_from >= 0) {
      const StreamOut& r = cur[so.residual_from + 1];
      const long long rel = out_abs - r.abs_base;
      if (!r.raw_row_bytes || rel < 0 || rel + out_rows > r.rows) { err = "internal: streaming skip rows unavailable at step " + std::to_string(k); return false; }
      res = raw_ptr(r, rel);
      sg.res_pitch = r.cap;
    }
    if (st.kind == 4) {
      const Step& s1 = steps[k + 1];
      const ConvLayer& c1 = p->convs[s1.conv];
      RuArgs ra;
      ra.a = reinterpret_cast<const __nv_bfloat16*>(act_ptr(in, 0));
      ra.x = res;
      ra.stream_f16 = sf16;
      ra.w7 = c.w_umma; ra.w1 = c1.w_umma;
      ra.bias7 = c.bias;
      ra.s2_a = p->snakes[sp.epi_snake].a; ra.s2_inv_b = p->snakes[sp.epi_snake].inv_b;
      ra.bias1 = c1.bias;
      ra.out_raw = out.raw_row_bytes ? raw_ptr(out, out.rows) : nullptr;
      ra.out_act = out.act_row_bytes ? reinterpret_cast<__nv_bfloat16*>(act_ptr(out, out.rows)) : nullptr;
      ra.act_f16 = s1.act_f16 ? 1 : 0;
      if (s1.epi_snake >= 0) { ra.sn_a = p->snakes[s1.epi_snake].a; ra.sn_inv_b = p->snakes[s1.epi_snake].inv_b; }
      if (!prepare_conv_ru(ra, B, 0, c.g.dilation, P.ru[k], err, &sg)) return false;
    } else if (st.kind == 6 || st.kind == 3) {
      // decoder tail: reads the stream (raw) tensor of the previous step, writes the staging waveform [B, io, out_rows]
      const bool preact = p->tail_preact && st.kind == 6;    // the last unit wrote SnakeBeta(x) in fp16 (kvae.cu, finalize_steps)
      if (preact ? !in.act_row_bytes : !in.raw_row_bytes) { err = "internal: streaming tail needs the stream tensor"; return false; }
      if (out_rows > s->max_samples) { err = "streaming: waveform staging overflow"; return false; }
      const int tanh_out = (p->arch.final_tanh) ? 1 : 0;
      P.wo_cout = c.g.Cout;
      if (st.kind == 6) {
        PreparedRun::WaveOutTc& t = P.wo_tc;
        std::memset(&t.p, 0, sizeof(t.p));
        t.p.pro_a = p->snakes[sp.pre_snake].a; t.p.pro_inv_b = p->snakes[sp.pre_snake].inv_b; t.p.w = c.w_direct;
        t.p.T = static_cast<int>(out_rows); t.p.B = B; t.p.COUT = c.g.Cout; t.p.tanh_out = tanh_out;
        t.p.tiles_per_clip = (t.p.T + kWoTcTile - 1) / kWoTcTile;
        t.p.total_tiles = t.p.tiles_per_clip * B;
        t.p.in_row0 = 0;                                   // -3 + bias 3
        t.p.y = s->wav_stage; t.p.y_f32 = 1;
        t.p.preact = preact ? 1 : 0;
        if (sf16) { if (!make_act_tmap(&t.tmX, preact ? act_ptr(in, 0) : raw_ptr(in, 0), B, static_cast<int>(in.rows), 128, 1, kWoTcRows, err, in.cap)) return false; }
        else if (!make_out_tmap(&t.tmX, raw_ptr(in, 0), B, static_cast<int>(in.rows), 128, 1, 1, err, kWoTcRows, in.cap)) return false;
      } else {
        WaveOutParams& w = P.wo_cc;
        std::memset(&w, 0, sizeof(w));
        w.x = reinterpret_cast<const float*>(raw_ptr(in, 0));
        w.pro_a = p->snakes[sp.pre_snake].a; w.pro_inv_b = p->snakes[sp.pre_snake].inv_b; w.w = c.w_direct;
        w.y = s->wav_stage; w.y_f32 = 1;
        w.T = static_cast<int>(out_rows); w.Cin = c.g.Cin; w.tanh_out = tanh_out;
        w.precise = (p->precision == KVAE_PREC_F32) ? 1 : 0;
        w.T_in = static_cast<int>(in.rows); w.in_row0 = 0; w.x_pitch = in.cap;
      }
      P.n_samples = out_rows;
    } else {
      ConvEpilogue ep;
      ep.bias = c.has_bias ? c.bias : nullptr;
      ep.residual = res;
      ep.residual_f32 = sf16 ? 0 : 1;
      ep.stream_f16 = sf16;
      ep.out_raw = out.raw_row_bytes ? raw_ptr(out, out.rows) : nullptr;
      ep.out_raw_f32 = sf16 ? 0 : 1;
      ep.out_act = out.act_row_bytes ? reinterpret_cast<__nv_bfloat16*>(act_ptr(out, out.rows)) : nullptr;
      if (sp.epi_snake >= 0) { ep.snake_a = p->snakes[sp.epi_snake].a; ep.snake_inv_b = p->snakes[sp.epi_snake].inv_b; }
      if (!c.umma) { ep.split3 = 1; ep.act_split = 1; ep.precise = 1; }
      ConvTuning2 tune;
      if (!prepare_conv_umma2(c.g, reinterpret_cast<const __nv_bfloat16*>(act_ptr(in, 0)), B, 0, c.w_umma, ep, tune,
                              P.umma2[k], err, &sg))
        return false;
    }
    P.launch[k] = 1;
    if (!last) out.rows += out_rows;
    (void)split;
  }
  // carry: keep the last `keep` rows of every buffer (all of them while a layer is still starved)
  for (size_t j = 0; j < cur.size(); ++j) {
    StreamOut& o = cur[j];
    if (!o.cap) continue;
    const long long keep = flush ? 0 : std::min<long long>(o.rows, o.keep);
    const long long drop = o.rows - keep;
    if (drop > 0 && keep > 0) {
      for (int which = 0; which < 2; ++which) {
        const int rb = which ? o.act_row_bytes : o.raw_row_bytes;
        if (!rb) continue;
        uint8_t* b0 = base + (which ? o.off_act : o.off_raw);
        StreamCarry cdesc{b0, b0 + static_cast<size_t>(drop) * rb, static_cast<unsigned long long>(keep) * rb,
                          static_cast<unsigned long long>(o.cap) * rb};
        if (cdesc.bytes % 16) { err = "internal: carry size not a multiple of 16 bytes"; return false; }
        (keep <= drop ? P.carry : P.carry_overlap).push_back(cdesc);
      }
    }
    o.abs_base += drop;
    o.rows = keep;
  }
  for (const StreamCarry& c : P.carry_overlap)
    if (c.bytes * static_cast<size_t>(B) > s->scratch_bytes) { err = "internal: carry scratch too small"; return false; }
  if (!P.carry.empty()) {
    if (cudaMalloc(&P.carry_dev, P.carry.size() * sizeof(StreamCarry)) != cudaSuccess) { err = "cudaMalloc(carry table) failed"; return false; }
    if (cudaMemcpy(P.carry_dev, P.carry.data(), P.carry.size() * sizeof(StreamCarry), cudaMemcpyHostToDevice) != cudaSuccess) {
      err = "cudaMemcpy(carry table) failed";
      return false;
    }
  }
  P.after = cur;
  (void)z_dtype;
  return true;
}

cudaError_t stream_launch_all(kvae_stream* s, PreparedPush& P, int n_frames, int z_f32, cudaStream_t st) {
  kvae_plan* p = s->plan;
  const int n = static_cast<int>(p->steps.size());
  const ConvLayer& c0 = p->convs[p->steps[0].conv];
  if (n_frames > 0) {
    // [B, D, n_frames] (staging) -> channels-last rows appended to the input window
    const StreamOut& o = s->outs[0];
    dim3 grid(ceil_div(n_frames, 32), ceil_div(c0.g.Cin, 32), s->B), block(32, 8);
    cf_to_cl_bf16_kernel<<<grid, block, 0, st>>>(s->z_stage, z_f32,
                                                 reinterpret_cast<__nv_bfloat16*>(s->ws + o.off_act + static_cast<size_t>(P.in_rows0) * o.act_row_bytes),
                                                 c0.g.Cin, n_frames, c0.umma ? 0 : 1, o.cap);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ++g_launches;
  }
  for (int k = 0; k < n; ++k) {
    if (!P.launch[k]) continue;
    const int kind = s->ss[k].kind;
    cudaError_t e = cudaSuccess;
    if (kind == 4) {
      e = launch_conv_ru(P.ru[k], st);
    } else if (kind == 0) {
      e = launch_conv_umma2(P.umma2[k], st);
    } else if (kind == 6) {
      static bool attr_set[64] = {false};
      int dev = 0;
      cudaGetDevice(&dev);
      if (!attr_set[dev & 63]) {
        e = cudaFuncSetAttribute(conv_wave_out_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wave_out_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        attr_set[dev & 63] = true;
      }
      const int grid = std::min(P.wo_tc.p.total_tiles, sm_count());
      if (s->stream_f16) e = launch_pdl(conv_wave_out_tc_kernel<true>, dim3(grid), dim3(kWoTcThreads), wave_out_tc_smem<true>(), st, P.wo_tc.tmX, P.wo_tc.p);
      else e = launch_pdl(conv_wave_out_tc_kernel<false>, dim3(grid), dim3(kWoTcThreads), wave_out_tc_smem<false>(), st, P.wo_tc.tmX, P.wo_tc.p);
    } else if (kind == 3) {
      e = launch_wave_out(P.wo_cc, P.wo_cout, s->B, st);
    }
    if (e != cudaSuccess) return e;
    ++g_launches;
  }
  if (!P.carry.empty()) {
    size_t maxb = 0;
    for (const StreamCarry& c : P.carry) maxb = std::max<size_t>(maxb, c.bytes);
    const int bx = static_cast<int>(std::max<size_t>(1, std::min<size_t>((maxb / 16 + 255) / 256, 8)));
    stream_carry_kernel<<<dim3(bx, static_cast<unsigned>(P.carry.size()), s->B), 256, 0, st>>>(P.carry_dev, static_cast<int>(P.carry.size()));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    ++g_launches;
  }
  for (const StreamCarry& c : P.carry_overlap) {           // rare (pushes shorter than a layer's halo): through scratch
    cudaError_t e = cudaMemcpy2DAsync(s->scratch, c.bytes, c.src, c.pitch, c.bytes, s->B, cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(c.dst, c.pitch, s->scratch, c.bytes, c.bytes, s->B, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

void stream_zero_initial_rows(kvae_stream* s, cudaStream_t st);

PreparedPush* stream_get(kvae_stream* s, int n_frames, bool flush, int z_dtype) {
  std::vector<long long> key;
  key.push_back(n_frames); key.push_back(flush ? 1 : 0); key.push_back(z_dtype);
  for (const StreamOut& o : s->outs) key.push_back(o.rows);
  auto it = s->cache.find(key);
  if (it == s->cache.end()) {
    auto P = std::make_unique<PreparedPush>();
    std::string err;
    if (!stream_prepare(s, n_frames, flush, z_dtype, *P, err)) { fail(err); return nullptr; }
    if (s->cache.size() > 64) s->cache.clear();
    it = s->cache.emplace(key, std::move(P)).first;
  }
  return it->second.get();
}

int stream_run(kvae_stream* s, const void* z, int z_dtype, int n_frames, bool flush, void* wav, int wav_dtype,
               long long wav_capacity, long long* n_samples, cudaStream_t st) {
  kvae_plan* p = s->plan;
  for (const ConvLayer& c : p->convs)
    if (!c.set) return fail("plan weights not set (kvae_plan_set_conv / kvae_plan_load_params)");
  for (const SnakeLayer& sn : p->snakes)
    if (!sn.set) return fail("plan SnakeBeta parameters not set");
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail("cannot select device");
  if (n_frames < 0 || n_frames > s->max_frames) return fail("streaming: n_frames out of range (0 .. max_frames)");
  PreparedPush* Pp = stream_get(s, n_frames, flush, z_dtype);
  if (!Pp) return -1;
  PreparedPush& P = *Pp;
  if (s->needs_zero) {
    stream_zero_initial_rows(s, st);
    s->needs_zero = false;
  }
  if (P.n_samples > wav_capacity) return fail("wav buffer too small for this call (kvae_decode_stream_samples tells the size)");
  const int io = p->arch.io_channels;
  if (n_frames > 0) {
    const size_t zb = static_cast<size_t>(s->B) * p->arch.latent_dim * n_frames * (z_dtype == KVAE_F32 ? 4 : 2);
    KV_CUDA(cudaMemcpyAsync(s->z_stage, z, zb, cudaMemcpyDeviceToDevice, st));
  }
  if (s->use_graphs) {
    if (!P.graph) {
      // the first call with this geometry runs eagerly (sets kernel attributes, fills caches), the second is captured
      // (captured on a stream of our own: the caller's may be the legacy default stream, which cannot capture)
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (st) KV_CUDA(cudaStreamIsCapturing(st, &cs));
      if (cs == cudaStreamCaptureStatusNone && P.eager_done && P.carry_overlap.empty()) {
        if (!s->cap_stream) KV_CUDA(cudaStreamCreateWithFlags(&s->cap_stream, cudaStreamNonBlocking));
        cudaGraph_t g = nullptr;
        KV_CUDA(cudaStreamBeginCapture(s->cap_stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t e = stream_launch_all(s, P, n_frames, z_dtype == KVAE_F32, s->cap_stream);
        cudaError_t e2 = cudaStreamEndCapture(s->cap_stream, &g);
        if (e != cudaSuccess) { if (g) cudaGraphDestroy(g); return cuda_fail(e, "stream capture"); }
        KV_CUDA(e2);
        e = cudaGraphInstantiate(&P.graph, g, 0);
        cudaGraphDestroy(g);
        KV_CUDA(e);
      }
    }
    if (P.graph) {
      KV_CUDA(cudaGraphLaunch(P.graph, st));
    } else {
      KV_CUDA(stream_launch_all(s, P, n_frames, z_dtype == KVAE_F32, st));
      P.eager_done = true;
    }
  } else {
    KV_CUDA(stream_launch_all(s, P, n_frames, z_dtype == KVAE_F32, st));
  }
  if (P.n_samples > 0) {
    // staging [B, io, n] fp32 -> the caller's buffer [B, io, n] (dtype conversion included)
    const size_t cnt = static_cast<size_t>(s->B) * io * P.n_samples;
    if (wav_dtype == KVAE_F32) {
      KV_CUDA(cudaMemcpyAsync(wav, s->wav_stage, cnt * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      const int blocks = static_cast<int>(std::min<size_t>((cnt + 255) / 256, 148 * 8));
      f32_to_bf16_kernel<<<blocks, 256, 0, st>>>(s->wav_stage, static_cast<__nv_bfloat16*>(wav), cnt);
      KV_CUDA(cudaGetLastError());
      ++g_launches;
    }
  }
  s->outs = P.after;
  s->frames_in += n_frames;
  s->samples_out += P.n_samples;
  if (n_samples) *n_samples = P.n_samples;
  if (flush) {
    stream_reset(s);
    s->needs_zero = true;
  }
  return 0;
}

// decides which kernels serve the steps and sizes the persistent window buffers
int stream_create(kvae_plan* p, int B, int max_frames, int use_graphs, kvae_stream** out) {
  if (!p || !out) return fail("null argument");
  if (p->direction != KVAE_DECODER) return fail("streaming: plan is not a decoder");
  if (B <= 0 || max_frames <= 0) return fail("streaming: bad batch / max_frames");
  DeviceGuard guard(p->device);
  if (!guard.ok) return fail("cannot select device");
  const std::vector<Step>& steps = p->steps;
  const int n = static_cast<int>(steps.size());
  auto s = std::make_unique<kvae_stream>();
  s->plan = p; s->B = B; s->max_frames = max_frames; s->use_graphs = use_graphs != 0;
  s->stream_f16 = p->stream_f16;
  s->ratio = p->ratio;
  s->ss.resize(n);
  s->outs.assign(n + 1, StreamOut());
  std::string err;
  const int split = (p->precision == KVAE_PREC_F32) ? 2 : 1;
  const ConvLayer& c0 = p->convs[steps[0].conv];
  if (!conv_tc(c0, false)) return fail("streaming: needs a tensor-core first layer (latent_dim and channels multiples of 64)");
  // step kinds and geometry
  for (int k = 0; k < n; ++k) {
    const Step& sp = steps[k];
    const ConvLayer& c = p->convs[sp.conv];
    StreamStep& st = s->ss[k];
    st.in_out = k - 1;
    if (sp.fuse == 2) { st.kind = 5; continue; }
    if (!stream_step_geometry(c.g, st, err)) return fail(err);
    if (sp.fuse == 1) st.kind = 4;
    else if (is_wave_out_step(p, steps, k)) st.kind = (p->precision == KVAE_PREC_BF16 && s->stream_f16) ? 6 : 3;
    else if (conv_tc(c, false)) st.kind = 0;
    else return fail("streaming: step " + std::to_string(k) + " has no tensor-core kernel (channels must be multiples of 64; "
                     "the 128-channel tail is required)");
    if (st.kind == 3 && (c.g.Cin != 128)) return fail("streaming: CUDA-core tail needs 128 input channels");
    if (k == n - 1 && st.kind == 0) return fail("streaming: tensor-core output conv unsupported");
  }
  // rate of every tensor relative to the latent frame rate, keep / init rows from the consumer
  std::vector<double> rates(n + 1, 1.0);
  for (int k = 0; k < n; ++k) rates[k + 1] = static_cast<double>(step_len(steps[k], 1000)) / 1000.0;
  size_t off = 0;
  auto consumer_of = [&](int j) -> int {   // first step that reads StreamOut j (j = -1: the input)
    for (int k = j + 1; k < n; ++k)
      if (s->ss[k].kind != 5 && s->ss[k].in_out == j) return k;
    return -1;
  };
  // a fused unit at step k writes the tensors of step k+1; its consumer is step k+2 reading StreamOut k+1
  for (int k = 0; k < n; ++k)
    if (s->ss[k].kind != 5 && k > 0 && s->ss[k - 1].kind == 5) s->ss[k].in_out = k - 1;
  for (int j = -1; j < n - 1; ++j) {
    StreamOut& o = s->outs[j + 1];
    if (j >= 0 && s->ss[j].kind == 4) continue;          // the k7 half of a fused unit produces no tensor
    const int cons = consumer_of(j);
    if (cons < 0) return fail("internal: streaming tensor without consumer");
    o.keep = s->ss[cons].span;
    o.init_rows = s->ss[cons].bias;
    const double r = rates[j + 1];
    o.cap = static_cast<long long>(std::ceil(r * (max_frames + 24))) + o.keep + 64;
    o.cap = (o.cap + 7) / 8 * 8;
    const int C = (j < 0) ? c0.g.Cin : p->convs[steps[j].conv].g.Cout;
    const bool needs_raw = j >= 0 && steps[j].needs_raw;
    const bool needs_act = j < 0 || steps[j].needs_act;
    if (needs_raw) {
      o.raw_row_bytes = C * (s->stream_f16 ? 2 : 4);
      o.off_raw = off;
      off += align_up(static_cast<size_t>(B) * o.cap * o.raw_row_bytes, 1024);
    }
    if (needs_act) {
      o.act_row_bytes = C * 2 * split;
      o.off_act = off;
      off += align_up(static_cast<size_t>(B) * o.cap * o.act_row_bytes, 1024);
    }
  }
  s->ws_bytes = std::max<size_t>(off, 1024);
  s->max_samples = static_cast<long long>(max_frames + 24) * p->ratio;
  size_t max_carry = 1024;
  for (const StreamOut& o : s->outs)
    max_carry = std::max<size_t>(max_carry, static_cast<size_t>(o.keep) * std::max(o.raw_row_bytes, o.act_row_bytes));
  s->scratch_bytes = max_carry * B;
  KV_CUDA(cudaMalloc(&s->ws, s->ws_bytes));
  KV_CUDA(cudaMemset(s->ws, 0, s->ws_bytes));            // the initial carry rows are zeros (conv left padding)
  KV_CUDA(cudaMalloc(&s->z_stage, static_cast<size_t>(B) * p->arch.latent_dim * max_frames * 4));
  KV_CUDA(cudaMalloc(&s->wav_stage, static_cast<size_t>(B) * p->arch.io_channels * s->max_samples * 4));
  KV_CUDA(cudaMalloc(&s->scratch, s->scratch_bytes));
  stream_reset(s.get());
  *out = s.release();
  return 0;
}

void stream_zero_initial_rows(kvae_stream* s, cudaStream_t st) {
  // the carry region of every buffer becomes the conv's left zero padding again
  for (const StreamOut& o : s->outs) {
    if (!o.cap || !o.init_rows) continue;
    if (o.raw_row_bytes)
      cudaMemset2DAsync(s->ws + o.off_raw, static_cast<size_t>(o.cap) * o.raw_row_bytes, 0, static_cast<size_t>(o.init_rows) * o.raw_row_bytes, s->B, st);
    if (o.act_row_bytes)
      cudaMemset2DAsync(s->ws + o.off_act, static_cast<size_t>(o.cap) * o.act_row_bytes, 0, static_cast<size_t>(o.init_rows) * o.act_row_bytes, s->B, st);
  }
}
