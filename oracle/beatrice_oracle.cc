// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into or called by the product
// (beatrice_vst_b200/csrc).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.
//
// PARITY UNPINNED: the reference's arithmetic for this path lives in a closed-source
// static library ("beatricelib" rc.0, pinned by URL in reference Makefile:24-28 and
// CMakeLists.txt:125-134) that is absent from /root/reference together with its
// weights, and the reference holds no test, golden vector or fixture at this boundary
// (SURVEY.md section 8c).  What this file restates is therefore the builder-defined
// network spec "M0" (beatrice_vst_b200/model_spec.py) behind the
// reference's exact C ABI, lib/beatricelib/beatrice.h:39-343.  It is pinned instead by
//   (1) an independent whole-utterance PyTorch model of the same spec
//       (oracle/torch_model.py, tests/test_oracle_vs_torch.py), and
//   (2) the reference's own unmodified call site src/common/*.cc linked against it
//       (oracle/Makefile -> oracle/_ref/, tests/test_callsite.py).
//
// Plain scalar fp32 C++ (the compiler may vectorise the channel loops); one stream per
// context, streaming one 10 ms frame per call like the reference ABI
// (beatrice.h:243-247, :266-271, :301-307).  State is kept as sliding windows that are
// shifted with memmove every frame -- deliberately a different mechanism from the ring
// buffers of the CUDA engine so the two implementations do not share a bug.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace oracle {

// ---- error codes: numeric values are contract (beatrice.h:30-36) ----
enum Err { kOk = 0, kOpen = 1, kTooSmall = 2, kTooLarge = 3, kInvalidSize = 4 };

constexpr uint32_t kMagic = 0x42323042u;
constexpr int kInHop = 160;    // beatrice.h:10
constexpr int kOutHop = 240;   // beatrice.h:11
constexpr int kHidden = 256;   // beatrice.h:13
constexpr int kCodebook = 512; // beatrice.h:25
constexpr int kKvLen = 384;    // beatrice.h:26
constexpr int kKvCh = 128;     // beatrice.h:27
constexpr int kBlocks = 4;     // beatrice.h:28
constexpr int kFeats = 4;      // beatrice.h:270 "output_pitch_feature // 4"
constexpr int kFormants = 9;   // beatrice.h:284-285

struct Dims {
  int family;          // 0 = 20a2, 1 = 20b1, 2 = 20rc0
  int phone_channels;  // beatrice.h:17,20,23
  int pitch_bins;      // beatrice.h:18,21,24
  bool has_setter;     // rc0 only (beatrice.h:207-209)
};
constexpr Dims kDims[3] = {{0, 256, 384, false}, {1, 256, 384, false}, {2, 128, 448, true}};

// ---------------------------------------------------------------------------------
// file reading
// ---------------------------------------------------------------------------------
struct Blob {
  uint32_t family = 0, kind = 0, count = 0;
  std::vector<float> data;
};

// Reads header + payload.  expected_floats < 0 means "derive from header count via
// floats_for_count".  Error mapping: cannot open -> kOpen; shorter than expected ->
// kTooSmall; longer -> kTooLarge; malformed header / not a multiple of 4 -> kInvalidSize.
template <class F>
static Err ReadBlob(const char* filename, uint32_t family, uint32_t kind_a, uint32_t kind_b,
                    F floats_for_count, Blob* out, bool header_only = false) {
  FILE* f = std::fopen(filename, "rb");
  if (!f) return kOpen;
  std::fseek(f, 0, SEEK_END);
  const long size = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  uint32_t h[4];
  if (size < 16) {
    std::fclose(f);
    return kTooSmall;
  }
  if (std::fread(h, 4, 4, f) != 4) {
    std::fclose(f);
    return kOpen;
  }
  if (h[0] != kMagic || h[1] != family || (h[2] != kind_a && h[2] != kind_b) ||
      (size - 16) % 4 != 0) {
    std::fclose(f);
    return kInvalidSize;
  }
  const long expect = floats_for_count(h[3]);
  const long have = (size - 16) / 4;
  if (expect < 0) {
    std::fclose(f);
    return kInvalidSize;
  }
  if (have < expect) {
    std::fclose(f);
    return kTooSmall;
  }
  if (have > expect) {
    std::fclose(f);
    return kTooLarge;
  }
  out->family = h[1];
  out->kind = h[2];
  out->count = h[3];
  if (!header_only) {
    out->data.resize(have);
    if (have > 0 && std::fread(out->data.data(), 4, have, f) != static_cast<size_t>(have)) {
      std::fclose(f);
      return kOpen;
    }
  }
  std::fclose(f);
  return kOk;
}

// ---------------------------------------------------------------------------------
// elementary ops
// ---------------------------------------------------------------------------------
static inline float Gelu(float x) { return 0.5f * x * (1.0f + std::erf(x * 0.70710678118654752f)); }
static inline float Lrelu(float x) { return x > 0.0f ? x : 0.1f * x; }

// A causal Conv1d with weights w[k][cin][cout], bias b[cout], stride s, dilation d.
// Output step t of a frame reads input steps u = t*s + (s-1) - (k-1-j)*d, j = 0..k-1,
// relative to the first new input step of the frame; u < 0 reaches into history.
struct ConvSpec {
  int k = 0, cin = 0, cout = 0, stride = 1, dil = 1;
  const float* w = nullptr;
  const float* b = nullptr;
  int History() const { return (k - 1) * dil - (stride - 1); }
};

// Sliding window holding `hist` history rows followed by the `t_in` rows of the
// current frame, channel-last.
struct Window {
  int hist = 0, t_in = 0, c = 0;
  std::vector<float> buf;
  void Init(int hist_rows, int rows_per_frame, int channels) {
    hist = hist_rows;
    t_in = rows_per_frame;
    c = channels;
    buf.assign(static_cast<size_t>(hist + t_in) * c, 0.0f);
  }
  // Slide by one frame and return where the new rows go.
  float* Advance() {
    if (hist > 0) std::memmove(buf.data(), buf.data() + static_cast<size_t>(t_in) * c,
                               static_cast<size_t>(hist) * c * sizeof(float));
    return buf.data() + static_cast<size_t>(hist) * c;
  }
  const float* Row(int u) const { return buf.data() + static_cast<size_t>(hist + u) * c; }
  float* Cur() { return buf.data() + static_cast<size_t>(hist) * c; }
};

// out[t_out][cout] (row stride = cout) from a window.  `win.hist` may exceed the conv's
// own history (shared windows use the max over their consumers).
static void RunConv(const ConvSpec& cv, const Window& win, int t_out, float* out) {
  const int co_n = cv.cout;
  for (int t = 0; t < t_out; ++t) {
    float* acc = out + static_cast<size_t>(t) * co_n;
    for (int co = 0; co < co_n; ++co) acc[co] = cv.b ? cv.b[co] : 0.0f;
    for (int j = 0; j < cv.k; ++j) {
      const int u = t * cv.stride + (cv.stride - 1) - (cv.k - 1 - j) * cv.dil;
      const float* x = win.Row(u);
      const float* wj = cv.w + static_cast<size_t>(j) * cv.cin * co_n;
      for (int ci = 0; ci < cv.cin; ++ci) {
        const float xv = x[ci];
        const float* wr = wj + static_cast<size_t>(ci) * co_n;
        for (int co = 0; co < co_n; ++co) acc[co] += xv * wr[co];
      }
    }
  }
}

// Takes consecutive tensors out of a payload in file order.
struct Cursor {
  const float* p;
  const float* end;
  const float* Take(size_t n) {
    const float* r = p;
    p += n;
    return r;
  }
};

static ConvSpec TakeConv(Cursor* c, int k, int cin, int cout, int stride, int dil) {
  ConvSpec s;
  s.k = k;
  s.cin = cin;
  s.cout = cout;
  s.stride = stride;
  s.dil = dil;
  s.w = c->Take(static_cast<size_t>(k) * cin * cout);
  s.b = c->Take(cout);
  return s;
}

// ---------------------------------------------------------------------------------
// strided-conv front end + normalised residual backbone (PhoneExtractor and
// PitchEstimator share the topology at different widths)
// ---------------------------------------------------------------------------------
struct FrontSpec {
  int k, cin, cout, stride;
};
static const FrontSpec kPhoneFront[6] = {{10, 1, 32, 5},   {3, 32, 64, 2},   {3, 64, 128, 2},
                                         {3, 128, 256, 2}, {3, 256, 256, 2}, {2, 256, 256, 2}};
static const int kPhoneDil[6] = {1, 2, 4, 1, 2, 4};
static const FrontSpec kPitchFront[6] = {{10, 1, 16, 5},  {3, 16, 32, 2},   {3, 32, 64, 2},
                                         {3, 64, 128, 2}, {3, 128, 128, 2}, {2, 128, 128, 2}};
static const int kPitchDil[3] = {1, 2, 4};

struct EncoderWeights {
  std::vector<float> payload;
  ConvSpec front[6];
  int n_res = 0, width = 0, head_out = 0;
  std::vector<const float*> gamma, beta;
  std::vector<ConvSpec> res;
  ConvSpec head;
  bool loaded = false;

  static size_t Count(const FrontSpec* fs, int n_res, int width, int head_out) {
    size_t n = 0;
    for (int i = 0; i < 6; ++i) n += static_cast<size_t>(fs[i].k) * fs[i].cin * fs[i].cout + fs[i].cout;
    n += static_cast<size_t>(n_res) * (2 * width + 3 * width * width + width);
    n += static_cast<size_t>(width) * head_out + head_out;
    return n;
  }
  void Bind(const FrontSpec* fs, const int* dil, int n_res_, int width_, int head_out_) {
    n_res = n_res_;
    width = width_;
    head_out = head_out_;
    Cursor c{payload.data(), payload.data() + payload.size()};
    for (int i = 0; i < 6; ++i) front[i] = TakeConv(&c, fs[i].k, fs[i].cin, fs[i].cout, fs[i].stride, 1);
    gamma.resize(n_res);
    beta.resize(n_res);
    res.resize(n_res);
    for (int i = 0; i < n_res; ++i) {
      gamma[i] = c.Take(width);
      beta[i] = c.Take(width);
      res[i] = TakeConv(&c, 3, width, width, 1, dil[i]);
    }
    head = TakeConv(&c, 1, width, head_out, 1, 1);
    loaded = true;
  }
};

struct EncoderState {
  Window front_in[6];           // input windows of the six strided convs
  std::vector<Window> res_in;   // windows of GELU(ChanNorm(x)) feeding each residual conv
  std::vector<float> scratch_a, scratch_b;
  bool ready = false;
  void Init(const EncoderWeights& w) {
    int t = kInHop;
    for (int i = 0; i < 6; ++i) {
      front_in[i].Init(w.front[i].History(), t, w.front[i].cin);
      t /= w.front[i].stride;
    }
    res_in.resize(w.n_res);
    for (int i = 0; i < w.n_res; ++i) res_in[i].Init(w.res[i].History(), 1, w.width);
    scratch_a.assign(kInHop * 64, 0.0f);
    scratch_b.assign(kInHop * 64, 0.0f);
    ready = true;
  }
};

// One 10 ms frame through front end, backbone and 1x1 head.  out has head_out floats.
static void RunEncoder(const EncoderWeights& w, EncoderState* st, const float* in160, float* out) {
  if (!st->ready) st->Init(w);
  int t = kInHop;
  std::memcpy(st->front_in[0].Advance(), in160, sizeof(float) * kInHop);
  std::vector<float>& tmp = st->scratch_a;
  for (int i = 0; i < 6; ++i) {
    const ConvSpec& cv = w.front[i];
    const int t_out = t / cv.stride;
    tmp.resize(static_cast<size_t>(t_out) * cv.cout);
    RunConv(cv, st->front_in[i], t_out, tmp.data());
    for (float& v : tmp) v = Gelu(v);
    if (i + 1 < 6) std::memcpy(st->front_in[i + 1].Advance(), tmp.data(), tmp.size() * sizeof(float));
    t = t_out;
  }
  // tmp now holds x[width] (one row per frame)
  const int c = w.width;
  std::vector<float> x(tmp.begin(), tmp.begin() + c);
  std::vector<float> y(c);
  for (int i = 0; i < w.n_res; ++i) {
    float mean = 0.0f;
    for (int ch = 0; ch < c; ++ch) mean += x[ch];
    mean /= static_cast<float>(c);
    float var = 0.0f;
    for (int ch = 0; ch < c; ++ch) {
      const float dlt = x[ch] - mean;
      var += dlt * dlt;
    }
    var /= static_cast<float>(c);
    const float rstd = 1.0f / std::sqrt(var + 1e-5f);
    float* g = st->res_in[i].Advance();
    for (int ch = 0; ch < c; ++ch) g[ch] = Gelu((x[ch] - mean) * rstd * w.gamma[i][ch] + w.beta[i][ch]);
    RunConv(w.res[i], st->res_in[i], 1, y.data());
    for (int ch = 0; ch < c; ++ch) x[ch] += y[ch];
  }
  Window head_in;
  head_in.Init(0, 1, c);
  std::memcpy(head_in.Cur(), x.data(), sizeof(float) * c);
  RunConv(w.head, head_in, 1, out);
}

// ---------------------------------------------------------------------------------
// model objects (immutable weights) and contexts (per-stream state); beatrice.h:211-227
// ---------------------------------------------------------------------------------
struct PhoneExtractor {
  Dims dims;
  EncoderWeights w;
};
struct PhoneContext {
  Dims dims;
  EncoderState st;
  int vq_neighbors = 0;
  const float* codebook = nullptr;  // caller-owned, one speaker: 512 x phone_channels
};
struct PitchEstimator {
  Dims dims;
  EncoderWeights w;
};
struct PitchContext {
  Dims dims;
  EncoderState st;
  int min_q = 1, max_q = 0;  // max_q filled in at creation (bins - 1)
};

struct MrfBranch {
  ConvSpec c1[3], c2[3];
};
struct WaveWeights {
  std::vector<float> payload;
  ConvSpec embed_phone;
  const float* pitch_emb = nullptr;  // [bins][256]
  const float* feat_proj = nullptr;  // [4][256]
  ConvSpec pre;
  ConvSpec ups[4];           // k = 2, cout = r * C_out, bias replicated per phase in ups_bias
  std::vector<float> ups_bias[4];
  MrfBranch mrf[4][3];
  ConvSpec post;
  bool loaded = false;
};
static const int kRates[4] = {5, 4, 4, 3};
static const int kStageCh[5] = {256, 128, 64, 32, 16};
static const int kMrfK[3] = {3, 7, 11};
static const int kMrfD[3] = {1, 3, 5};

static size_t WaveCount(const Dims& d) {
  size_t n = static_cast<size_t>(d.phone_channels) * kHidden + kHidden;
  n += static_cast<size_t>(d.pitch_bins) * kHidden + kFeats * kHidden;
  n += 7u * kHidden * kHidden + kHidden;
  for (int s = 0; s < 4; ++s) {
    const int cin = kStageCh[s], cout = kStageCh[s + 1];
    n += 2u * cin * kRates[s] * cout + cout;
    for (int k : kMrfK) n += 3u * 2u * (static_cast<size_t>(k) * cout * cout + cout);
  }
  n += 7u * 16 + 1;
  return n;
}

struct WaveformGenerator {
  Dims dims;
  WaveWeights w;
  void Bind() {
    Cursor c{w.payload.data(), w.payload.data() + w.payload.size()};
    w.embed_phone = TakeConv(&c, 1, dims.phone_channels, kHidden, 1, 1);
    w.pitch_emb = c.Take(static_cast<size_t>(dims.pitch_bins) * kHidden);
    w.feat_proj = c.Take(kFeats * kHidden);
    w.pre = TakeConv(&c, 7, kHidden, kHidden, 1, 1);
    for (int s = 0; s < 4; ++s) {
      const int cin = kStageCh[s], cout = kStageCh[s + 1], r = kRates[s];
      ConvSpec u;
      u.k = 2;
      u.cin = cin;
      u.cout = r * cout;
      u.w = c.Take(2u * cin * r * cout);
      const float* b = c.Take(cout);
      w.ups_bias[s].resize(static_cast<size_t>(r) * cout);
      for (int p = 0; p < r; ++p) std::copy(b, b + cout, w.ups_bias[s].begin() + static_cast<size_t>(p) * cout);
      u.b = w.ups_bias[s].data();
      w.ups[s] = u;
      for (int ki = 0; ki < 3; ++ki)
        for (int di = 0; di < 3; ++di) {
          w.mrf[s][ki].c1[di] = TakeConv(&c, kMrfK[ki], cout, cout, 1, kMrfD[di]);
          w.mrf[s][ki].c2[di] = TakeConv(&c, kMrfK[ki], cout, cout, 1, 1);
        }
    }
    w.post = TakeConv(&c, 7, 16, 1, 1, 1);
    w.loaded = true;
  }
};

struct WaveformContext {
  Dims dims;
  bool ready = false;
  Window pre_in;         // hidden rows feeding the k=7 pre conv
  Window ups_in[4];      // lrelu(x) rows feeding each upsampler (1 row of history)
  Window c1_in[4][3][3]; // lrelu(y) windows
  Window c2_in[4][3][3]; // lrelu(a) windows
  Window post_in;
  // conditioning written by the setters (beatrice.h:323-343); identity until set
  float spk_add[kHidden];
  float formant_add[kHidden];
  std::vector<float> film[4];  // [gamma(C) | beta(C)]
  // debug taps of the latest frame (tests only)
  std::vector<float> tap_hidden, tap_pre, tap_stage[4];
  WaveformContext() {
    std::fill(spk_add, spk_add + kHidden, 0.0f);
    std::fill(formant_add, formant_add + kHidden, 0.0f);
    for (int s = 0; s < 4; ++s) film[s].assign(2 * kStageCh[s + 1], 0.0f);
  }
  void Init(const WaveWeights& w) {
    pre_in.Init(w.pre.History(), 1, kHidden);
    int t = 1;
    for (int s = 0; s < 4; ++s) {
      ups_in[s].Init(1, t, kStageCh[s]);
      t *= kRates[s];
      const int c = kStageCh[s + 1];
      for (int ki = 0; ki < 3; ++ki)
        for (int di = 0; di < 3; ++di) {
          c1_in[s][ki][di].Init(w.mrf[s][ki].c1[di].History(), t, c);
          c2_in[s][ki][di].Init(w.mrf[s][ki].c2[di].History(), t, c);
        }
    }
    post_in.Init(w.post.History(), kOutHop, 16);
    ready = true;
  }
};

struct EmbeddingSetter {
  Dims dims;
  std::vector<float> payload;
  const float *add_w = nullptr, *add_b = nullptr, *for_w = nullptr, *for_b = nullptr;
  const float *query[4] = {}, *film_w[4] = {}, *film_b[4] = {};
  bool loaded = false;
  static size_t Count() {
    size_t n = 2u * (kHidden * kHidden + kHidden);
    for (int b = 0; b < kBlocks; ++b) n += kKvCh + static_cast<size_t>(kKvCh) * 2 * kStageCh[b + 1] + 2 * kStageCh[b + 1];
    return n;
  }
  void Bind() {
    Cursor c{payload.data(), payload.data() + payload.size()};
    add_w = c.Take(kHidden * kHidden);
    add_b = c.Take(kHidden);
    for_w = c.Take(kHidden * kHidden);
    for_b = c.Take(kHidden);
    for (int b = 0; b < kBlocks; ++b) {
      query[b] = c.Take(kKvCh);
      film_w[b] = c.Take(static_cast<size_t>(kKvCh) * 2 * kStageCh[b + 1]);
      film_b[b] = c.Take(2 * kStageCh[b + 1]);
    }
    loaded = true;
  }
};
struct EmbeddingContext {
  std::vector<float> kv;  // copy of the registered 384 x 128 embedding
  bool registered = false;
};

// ---------------------------------------------------------------------------------
// per-frame entry points
// ---------------------------------------------------------------------------------

// beatrice.h:243-247 (rc0), :65-69 (a2), :148-152 (b1)
static void ExtractPhone(const PhoneExtractor* pe, const float* in, float* out, PhoneContext* ctx) {
  if (!pe->w.loaded) {
    std::fill(out, out + pe->dims.phone_channels, 0.0f);
    return;
  }
  RunEncoder(pe->w, &ctx->st, in, out);
  // optional kNN-VQ against the current speaker codebook (beatrice.h:239-242, :318-322):
  // replace the feature by the mean of its n nearest (squared L2) codebook rows.
  const int n = ctx->vq_neighbors;
  const int c = pe->dims.phone_channels;
  if (pe->dims.has_setter && n > 0 && ctx->codebook != nullptr) {
    std::vector<float> dist(kCodebook);
    for (int i = 0; i < kCodebook; ++i) {
      const float* e = ctx->codebook + static_cast<size_t>(i) * c;
      float nn = 0.0f, dot = 0.0f;
      for (int ch = 0; ch < c; ++ch) {
        nn += e[ch] * e[ch];
        dot += e[ch] * out[ch];
      }
      dist[i] = nn - 2.0f * dot;
    }
    std::vector<float> acc(c, 0.0f);
    const int take = std::min(n, kCodebook);
    for (int r = 0; r < take; ++r) {
      int best = 0;
      for (int i = 1; i < kCodebook; ++i)
        if (dist[i] < dist[best]) best = i;
      dist[best] = INFINITY;
      const float* e = ctx->codebook + static_cast<size_t>(best) * c;
      for (int ch = 0; ch < c; ++ch) acc[ch] += e[ch];
    }
    const float inv = 1.0f / static_cast<float>(take);
    for (int ch = 0; ch < c; ++ch) out[ch] = acc[ch] * inv;
  }
}

// beatrice.h:266-271: arg-max of the bin logits restricted to [min,max] + 4 features
static void EstimatePitch(const PitchEstimator* pi, const float* in, int* q, float* feat, PitchContext* ctx) {
  const int bins = pi->dims.pitch_bins;
  if (!pi->w.loaded) {
    *q = 1;
    std::fill(feat, feat + kFeats, 0.0f);
    return;
  }
  std::vector<float> head(bins + kFeats);
  RunEncoder(pi->w, &ctx->st, in, head.data());
  int lo = std::clamp(ctx->min_q, 1, bins - 1);
  int hi = std::clamp(ctx->max_q, 1, bins - 1);
  if (hi < lo) hi = lo;
  int best = lo;
  for (int i = lo + 1; i <= hi; ++i)
    if (head[i] > head[best]) best = i;
  *q = best;
  for (int i = 0; i < kFeats; ++i) feat[i] = head[bins + i];
}

// beatrice.h:301-307 (rc0); :112-120 / :195-203 pass the speaker vector per call.
static void GenerateWaveform(const WaveformGenerator* wg, const float* phone, const int* qp,
                             const float* feat, const float* speaker_or_null, float* out,
                             WaveformContext* ctx) {
  const WaveWeights& w = wg->w;
  if (!w.loaded) {
    std::fill(out, out + kOutHop, 0.0f);
    return;
  }
  if (!ctx->ready) ctx->Init(w);
  const int bins = wg->dims.pitch_bins;
  const int q = std::clamp(*qp, 0, bins - 1);
  // conditioning: phone 1x1 + pitch-bin embedding + feature projection + speaker + formant
  float* h = ctx->pre_in.Advance();
  {
    Window ph;
    ph.Init(0, 1, wg->dims.phone_channels);
    std::memcpy(ph.Cur(), phone, sizeof(float) * wg->dims.phone_channels);
    RunConv(w.embed_phone, ph, 1, h);
    const float* pe = w.pitch_emb + static_cast<size_t>(q) * kHidden;
    for (int c = 0; c < kHidden; ++c) {
      float v = h[c] + pe[c];
      float fp = 0.0f;
      for (int i = 0; i < kFeats; ++i) fp += feat[i] * w.feat_proj[i * kHidden + c];
      v += fp;
      if (wg->dims.has_setter) {
        v += ctx->spk_add[c];
        v += ctx->formant_add[c];
      } else if (speaker_or_null) {
        v += speaker_or_null[c];
      }
      h[c] = v;
    }
  }
  ctx->tap_hidden.assign(h, h + kHidden);
  std::vector<float> x(kHidden);
  RunConv(w.pre, ctx->pre_in, 1, x.data());
  ctx->tap_pre = x;
  int t = 1;
  std::vector<float> u, a, y, sum;
  for (int s = 0; s < 4; ++s) {
    const int cin = kStageCh[s], c = kStageCh[s + 1], r = kRates[s];
    float* ui = ctx->ups_in[s].Advance();
    for (int i = 0; i < t * cin; ++i) ui[i] = Lrelu(x[i]);
    u.resize(static_cast<size_t>(t) * r * c);
    RunConv(w.ups[s], ctx->ups_in[s], t, u.data());  // [t][r*c] == [t*r][c]
    t *= r;
    if (wg->dims.has_setter) {
      const float* g = ctx->film[s].data();
      const float* bt = g + c;
      for (int row = 0; row < t; ++row)
        for (int ch = 0; ch < c; ++ch) {
          float& v = u[static_cast<size_t>(row) * c + ch];
          v = v * (1.0f + g[ch]) + bt[ch];
        }
    }
    const size_t n = static_cast<size_t>(t) * c;
    sum.assign(n, 0.0f);
    a.resize(n);
    for (int ki = 0; ki < 3; ++ki) {
      y = u;
      for (int di = 0; di < 3; ++di) {
        float* in1 = ctx->c1_in[s][ki][di].Advance();
        for (size_t i = 0; i < n; ++i) in1[i] = Lrelu(y[i]);
        RunConv(w.mrf[s][ki].c1[di], ctx->c1_in[s][ki][di], t, a.data());
        float* in2 = ctx->c2_in[s][ki][di].Advance();
        for (size_t i = 0; i < n; ++i) in2[i] = Lrelu(a[i]);
        RunConv(w.mrf[s][ki].c2[di], ctx->c2_in[s][ki][di], t, a.data());
        for (size_t i = 0; i < n; ++i) y[i] += a[i];
      }
      for (size_t i = 0; i < n; ++i) sum[i] += y[i];
    }
    x.resize(n);
    for (size_t i = 0; i < n; ++i) x[i] = sum[i] * (1.0f / 3.0f);
    ctx->tap_stage[s] = x;
  }
  float* pin = ctx->post_in.Advance();
  for (int i = 0; i < kOutHop * 16; ++i) pin[i] = Lrelu(x[i]);
  RunConv(w.post, ctx->post_in, kOutHop, out);
  for (int i = 0; i < kOutHop; ++i) out[i] = std::tanh(out[i]);
}

// y[256] = W^T e + b with W stored [in][out]
static void Project256(const float* w, const float* b, const float* e, float* y) {
  for (int o = 0; o < kHidden; ++o) y[o] = b[o];
  for (int i = 0; i < kHidden; ++i) {
    const float ev = e[i];
    const float* wr = w + static_cast<size_t>(i) * kHidden;
    for (int o = 0; o < kHidden; ++o) y[o] += ev * wr[o];
  }
}

// beatrice.h:339-343: attention-pool the registered 384x128 embedding with the block's
// query, then a linear layer gives the block's FiLM (gamma | beta).
static void SetKvBlock(const EmbeddingSetter* es, int block, const EmbeddingContext* ec, WaveformContext* wc) {
  if (!es->loaded || !ec->registered || block < 0 || block >= kBlocks) return;
  const int c = kStageCh[block + 1];
  std::vector<float> score(kKvLen);
  float mx = -INFINITY;
  const float scale = 1.0f / std::sqrt(static_cast<float>(kKvCh));
  for (int i = 0; i < kKvLen; ++i) {
    float s = 0.0f;
    for (int ch = 0; ch < kKvCh; ++ch) s += ec->kv[static_cast<size_t>(i) * kKvCh + ch] * es->query[block][ch];
    score[i] = s * scale;
    mx = std::max(mx, score[i]);
  }
  float den = 0.0f;
  for (int i = 0; i < kKvLen; ++i) {
    score[i] = std::exp(score[i] - mx);
    den += score[i];
  }
  std::vector<float> pooled(kKvCh, 0.0f);
  for (int i = 0; i < kKvLen; ++i) {
    const float p = score[i] / den;
    for (int ch = 0; ch < kKvCh; ++ch) pooled[ch] += p * ec->kv[static_cast<size_t>(i) * kKvCh + ch];
  }
  std::vector<float>& f = wc->film[block];
  for (int o = 0; o < 2 * c; ++o) f[o] = es->film_b[block][o];
  for (int i = 0; i < kKvCh; ++i)
    for (int o = 0; o < 2 * c; ++o) f[o] += pooled[i] * es->film_w[block][static_cast<size_t>(i) * 2 * c + o];
}

static size_t SpeakerFloats(const Dims& d, uint32_t n) {
  if (d.has_setter)
    return static_cast<size_t>(kFormants) * kHidden +
           static_cast<size_t>(n) * (static_cast<size_t>(kCodebook) * d.phone_channels + kHidden +
                                     static_cast<size_t>(kKvLen) * kKvCh);
  return static_cast<size_t>(n) * kHidden;
}

template <class Obj>
static int ReadEncoder(Obj* obj, const char* fn, uint32_t kind, const FrontSpec* fs, const int* dil,
                       int n_res, int width, int head_out) {
  const size_t expect = EncoderWeights::Count(fs, n_res, width, head_out);
  Blob b;
  const Err e = ReadBlob(fn, obj->dims.family, kind, kind,
                         [&](uint32_t cnt) { return cnt == expect ? static_cast<long>(expect) : -1L; }, &b);
  if (e != kOk) return e;
  obj->w.payload = std::move(b.data);
  obj->w.Bind(fs, dil, n_res, width, head_out);
  return kOk;
}

}  // namespace oracle

// =================================================================================
// C ABI -- one stamp per API family (beatrice.h:39-121, :122-203, :205-343)
// =================================================================================
using namespace oracle;

#define ORACLE_COMMON_API(PFX, FAM)                                                               \
  extern "C" {                                                                                    \
  void* PFX##_CreatePhoneExtractor(void) {                                                        \
    auto* p = new PhoneExtractor();                                                               \
    p->dims = kDims[FAM];                                                                         \
    return p;                                                                                     \
  }                                                                                               \
  void PFX##_DestroyPhoneExtractor(void* p) { delete static_cast<PhoneExtractor*>(p); }           \
  void* PFX##_CreatePhoneContext1(void) {                                                         \
    auto* p = new PhoneContext();                                                                 \
    p->dims = kDims[FAM];                                                                         \
    return p;                                                                                     \
  }                                                                                               \
  void PFX##_DestroyPhoneContext1(void* p) { delete static_cast<PhoneContext*>(p); }              \
  int PFX##_ReadPhoneExtractorParameters(void* pe, const char* fn) {                              \
    auto* p = static_cast<PhoneExtractor*>(pe);                                                   \
    return ReadEncoder(p, fn, 1, kPhoneFront, kPhoneDil, 6, 256, p->dims.phone_channels);         \
  }                                                                                               \
  void PFX##_ExtractPhone1(const void* pe, const float* in, float* out, void* ctx) {              \
    ExtractPhone(static_cast<const PhoneExtractor*>(pe), in, out, static_cast<PhoneContext*>(ctx)); \
  }                                                                                               \
  void* PFX##_CreatePitchEstimator(void) {                                                        \
    auto* p = new PitchEstimator();                                                               \
    p->dims = kDims[FAM];                                                                         \
    return p;                                                                                     \
  }                                                                                               \
  void PFX##_DestroyPitchEstimator(void* p) { delete static_cast<PitchEstimator*>(p); }           \
  void* PFX##_CreatePitchContext1(void) {                                                         \
    auto* p = new PitchContext();                                                                 \
    p->dims = kDims[FAM];                                                                         \
    p->max_q = kDims[FAM].pitch_bins - 1;                                                         \
    return p;                                                                                     \
  }                                                                                               \
  void PFX##_DestroyPitchContext1(void* p) { delete static_cast<PitchContext*>(p); }              \
  int PFX##_ReadPitchEstimatorParameters(void* pi, const char* fn) {                              \
    auto* p = static_cast<PitchEstimator*>(pi);                                                   \
    return ReadEncoder(p, fn, 2, kPitchFront, kPitchDil, 3, 128, p->dims.pitch_bins + kFeats);    \
  }                                                                                               \
  void PFX##_SetMinQuantizedPitch(void* ctx, int v) { static_cast<PitchContext*>(ctx)->min_q = v; } \
  void PFX##_SetMaxQuantizedPitch(void* ctx, int v) { static_cast<PitchContext*>(ctx)->max_q = v; } \
  void PFX##_EstimatePitch1(const void* pi, const float* in, int* q, float* feat, void* ctx) {    \
    EstimatePitch(static_cast<const PitchEstimator*>(pi), in, q, feat, static_cast<PitchContext*>(ctx)); \
  }                                                                                               \
  int PFX##_ReadNSpeakers(const char* fn, int* out) {                                             \
    Blob b;                                                                                       \
    const Err e = ReadBlob(fn, FAM, 5, 6,                                                         \
                           [&](uint32_t n) { return static_cast<long>(SpeakerFloats(kDims[FAM], n)); }, \
                           &b, true);                                                             \
    if (e != kOk) return e;                                                                       \
    *out = static_cast<int>(b.count);                                                             \
    return kOk;                                                                                   \
  }                                                                                               \
  void* PFX##_CreateWaveformGenerator(void) {                                                     \
    auto* p = new WaveformGenerator();                                                            \
    p->dims = kDims[FAM];                                                                         \
    return p;                                                                                     \
  }                                                                                               \
  void PFX##_DestroyWaveformGenerator(void* p) { delete static_cast<WaveformGenerator*>(p); }     \
  void* PFX##_CreateWaveformContext1(void) {                                                      \
    auto* p = new WaveformContext();                                                              \
    p->dims = kDims[FAM];                                                                         \
    return p;                                                                                     \
  }                                                                                               \
  void PFX##_DestroyWaveformContext1(void* p) { delete static_cast<WaveformContext*>(p); }        \
  int PFX##_ReadWaveformGeneratorParameters(void* wg, const char* fn) {                           \
    auto* p = static_cast<WaveformGenerator*>(wg);                                                \
    const size_t expect = WaveCount(p->dims);                                                     \
    Blob b;                                                                                       \
    const Err e = ReadBlob(fn, FAM, 3, 3,                                                         \
                           [&](uint32_t cnt) { return cnt == expect ? static_cast<long>(expect) : -1L; }, &b); \
    if (e != kOk) return e;                                                                       \
    p->w.payload = std::move(b.data);                                                             \
    p->Bind();                                                                                    \
    return kOk;                                                                                   \
  }                                                                                               \
  }

ORACLE_COMMON_API(Beatrice20a2, 0)
ORACLE_COMMON_API(Beatrice20b1, 1)
ORACLE_COMMON_API(Beatrice20rc0, 2)

// ---- a2 / b1: speaker vector per call (beatrice.h:98-101, :112-120, :181-184, :195-203) ----
#define ORACLE_LEGACY_API(PFX, FAM)                                                               \
  extern "C" {                                                                                    \
  int PFX##_ReadSpeakerEmbeddings(const char* fn, float* out) {                                   \
    Blob b;                                                                                       \
    const Err e = ReadBlob(fn, FAM, 5, 6, [&](uint32_t n) { return static_cast<long>(n) * kHidden; }, &b); \
    if (e != kOk) return e;                                                                       \
    std::memcpy(out, b.data.data(), b.data.size() * sizeof(float));                               \
    return kOk;                                                                                   \
  }                                                                                               \
  void PFX##_GenerateWaveform1(const void* wg, const float* phone, const int* q, const float* feat, \
                               const float* spk, float* out, void* ctx) {                         \
    GenerateWaveform(static_cast<const WaveformGenerator*>(wg), phone, q, feat, spk, out,         \
                     static_cast<WaveformContext*>(ctx));                                         \
  }                                                                                               \
  }

ORACLE_LEGACY_API(Beatrice20a2, 0)
ORACLE_LEGACY_API(Beatrice20b1, 1)

// ---- rc0 extras (beatrice.h:239-242, :276-290, :301-343) ----
extern "C" {

void Beatrice20rc0_SetVQNumNeighbors(void* ctx, int n) {
  static_cast<PhoneContext*>(ctx)->vq_neighbors = std::clamp(n, 0, kCodebook);
}

int Beatrice20rc0_ReadSpeakerEmbeddings(const char* fn, float* codebook, float* additive, float* formant,
                                        float* kv) {
  Blob b;
  const Err e = ReadBlob(fn, 2, 5, 5,
                         [&](uint32_t n) { return static_cast<long>(SpeakerFloats(kDims[2], n)); }, &b);
  if (e != kOk) return e;
  const float* p = b.data.data();
  std::memcpy(formant, p, sizeof(float) * kFormants * kHidden);
  p += kFormants * kHidden;
  const size_t cb = static_cast<size_t>(kCodebook) * kDims[2].phone_channels;
  const size_t kvn = static_cast<size_t>(kKvLen) * kKvCh;
  for (uint32_t i = 0; i < b.count; ++i) {
    std::memcpy(codebook + i * cb, p, sizeof(float) * cb);
    p += cb;
    std::memcpy(additive + static_cast<size_t>(i) * kHidden, p, sizeof(float) * kHidden);
    p += kHidden;
    std::memcpy(kv + i * kvn, p, sizeof(float) * kvn);
    p += kvn;
  }
  return kOk;
}

void Beatrice20rc0_GenerateWaveform1(const void* wg, const float* phone, const int* q, const float* feat,
                                     float* out, void* ctx) {
  GenerateWaveform(static_cast<const WaveformGenerator*>(wg), phone, q, feat, nullptr, out,
                   static_cast<WaveformContext*>(ctx));
}

void* Beatrice20rc0_CreateEmbeddingSetter(void) {
  auto* p = new EmbeddingSetter();
  p->dims = kDims[2];
  return p;
}
void Beatrice20rc0_DestroyEmbeddingSetter(void* p) { delete static_cast<EmbeddingSetter*>(p); }
void* Beatrice20rc0_CreateEmbeddingContext(void) { return new EmbeddingContext(); }
void Beatrice20rc0_DestroyEmbeddingContext(void* p) { delete static_cast<EmbeddingContext*>(p); }

int Beatrice20rc0_ReadEmbeddingSetterParameters(void* es, const char* fn) {
  auto* p = static_cast<EmbeddingSetter*>(es);
  const size_t expect = EmbeddingSetter::Count();
  Blob b;
  const Err e = ReadBlob(fn, 2, 4, 4,
                         [&](uint32_t cnt) { return cnt == expect ? static_cast<long>(expect) : -1L; }, &b);
  if (e != kOk) return e;
  p->payload = std::move(b.data);
  p->Bind();
  return kOk;
}

// The call sites pass ONE speaker's 512x128 slice (processor_core_2.cc:118-121, :447-450)
// although the header comment says n_speakers * ... (beatrice.h:320-321).
void Beatrice20rc0_SetCodebook(void* phone_ctx, const float* codebook) {
  static_cast<PhoneContext*>(phone_ctx)->codebook = codebook;
}

void Beatrice20rc0_SetAdditiveSpeakerEmbedding(const void* es, const float* emb, void* /*ectx*/, void* wctx) {
  const auto* s = static_cast<const EmbeddingSetter*>(es);
  if (!s->loaded) return;
  Project256(s->add_w, s->add_b, emb, static_cast<WaveformContext*>(wctx)->spk_add);
}

void Beatrice20rc0_SetFormantShiftEmbedding(const void* es, const float* emb, void* /*ectx*/, void* wctx) {
  const auto* s = static_cast<const EmbeddingSetter*>(es);
  if (!s->loaded) return;
  Project256(s->for_w, s->for_b, emb, static_cast<WaveformContext*>(wctx)->formant_add);
}

void Beatrice20rc0_RegisterKeyValueSpeakerEmbedding(const void* /*es*/, const float* kv, void* ectx) {
  auto* c = static_cast<EmbeddingContext*>(ectx);
  c->kv.assign(kv, kv + static_cast<size_t>(kKvLen) * kKvCh);
  c->registered = true;
}

void Beatrice20rc0_SetKeyValueSpeakerEmbedding(const void* es, int block, void* ectx, void* wctx) {
  SetKvBlock(static_cast<const EmbeddingSetter*>(es), block, static_cast<const EmbeddingContext*>(ectx),
             static_cast<WaveformContext*>(wctx));
}

// ---- test-only taps: activations of the latest GenerateWaveform1 call ----
// which: 0 hidden[256], 1 pre[256], 2..5 stage output [T*C]; returns the float count.
int BeatriceOracle_WaveformTap(const void* wctx, int which, float* out, int capacity) {
  const auto* c = static_cast<const WaveformContext*>(wctx);
  const std::vector<float>* v = nullptr;
  if (which == 0) v = &c->tap_hidden;
  else if (which == 1) v = &c->tap_pre;
  else if (which >= 2 && which < 6) v = &c->tap_stage[which - 2];
  if (!v) return -1;
  const int n = static_cast<int>(v->size());
  if (out && capacity >= n) std::memcpy(out, v->data(), sizeof(float) * n);
  return n;
}

}  // extern "C"
