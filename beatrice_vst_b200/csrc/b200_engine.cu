#include "b200_engine.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

namespace b200 {

std::atomic<uint64_t> g_kernel_launches{0};

// ---------------------------------------------------------------------------------------
// sticky failure state (b200_common.h): first failure wins, later per-frame calls write silence
// ---------------------------------------------------------------------------------------
namespace {
std::atomic<int> g_error_code{0};
std::mutex g_error_mu;
char g_error_text[512] = "";
bool AbortOnError() {
  static const bool on = [] {
    const char* e = std::getenv("BEATRICE_B200_ABORT_ON_ERROR");
    return e && e[0] == '1';
  }();
  return on;
}
void Latch(int code, const char* text) noexcept {
  std::lock_guard<std::mutex> lock(g_error_mu);
  if (g_error_code.load(std::memory_order_relaxed) != 0) return;
  std::snprintf(g_error_text, sizeof(g_error_text), "%s", text);
  std::fprintf(stderr, "[libbeatrice_b200] ERROR (latched; output is silence until cleared): %s\n", g_error_text);
  g_error_code.store(code, std::memory_order_release);
}
}  // namespace

// Dead-lock records of the tensor-core kernels (b200_tc_common.cuh, SpinTimeout): a small mapped pinned buffer per
// process, its device address installed in every translation unit that polls.
void SetSpinDebugMrf(unsigned long long*);
void SetSpinDebugMrfc(unsigned long long*);
void SetSpinDebugEnc(unsigned long long*);
namespace {
unsigned long long* g_spin_host = nullptr;
std::mutex g_spin_mu;
bool g_spin_installed[64] = {};
}  // namespace
void InstallSpinDebug(int device) {
  std::lock_guard<std::mutex> lock(g_spin_mu);
  if (device < 0 || device >= 64 || g_spin_installed[device]) return;
  if (!g_spin_host) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&g_spin_host), 4096, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      (void)cudaGetLastError();
      g_spin_host = nullptr;
      return;
    }
    std::memset(g_spin_host, 0, 4096);
  }
  unsigned long long* dev = nullptr;
  if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev), g_spin_host, 0) != cudaSuccess) {
    (void)cudaGetLastError();
    return;
  }
  SetSpinDebugMrf(dev);
  SetSpinDebugMrfc(dev);
  SetSpinDebugEnc(dev);
  (void)cudaGetLastError();
  g_spin_installed[device] = true;
}
static void PrintSpinRecords() {
  if (!g_spin_host || g_spin_host[0] == 0) return;
  const unsigned long long n = g_spin_host[0] < 60 ? g_spin_host[0] : 60;
  for (unsigned long long i = 0; i < n; ++i) {
    const unsigned long long* r = g_spin_host + 8 + i * 8;
    std::fprintf(stderr, "[libbeatrice_b200]   stalled wait: site %llu (file id * 1000 + line) block (%llu,%llu) have/parity %llu want/info %llu sm %llu thread %llu\n",
                 r[0], r[1] & 0xffffffffull, r[1] >> 32, r[2] & 0xffffffffull, r[2] >> 32, r[3], r[4]);
  }
}

void Fail(int code, const char* what, const char* file, int line) {
  char text[480];
  if (code > 0) {
    std::snprintf(text, sizeof(text), "%s:%d: %s -> %s", file, line, what, cudaGetErrorString(static_cast<cudaError_t>(code)));
    (void)cudaGetLastError();   // non-sticky CUDA errors are consumed; sticky ones keep failing and keep being caught
  } else {
    std::snprintf(text, sizeof(text), "%s:%d: %s", file, line, what);
  }
  Latch(code, text);
  PrintSpinRecords();
  if (AbortOnError()) std::abort();
  throw Failure{code};
}
bool Failed() { return g_error_code.load(std::memory_order_acquire) != 0; }
int LastErrorCode() { return g_error_code.load(std::memory_order_acquire); }
const char* LastErrorText() { return Failed() ? g_error_text : ""; }
void ClearError() {
  std::lock_guard<std::mutex> lock(g_error_mu);
  (void)cudaGetLastError();
  g_error_code.store(0, std::memory_order_release);
  g_error_text[0] = 0;
}
void NoteException(const char* what) noexcept {
  char text[480];
  std::snprintf(text, sizeof(text), "host exception: %s", what ? what : "?");
  Latch(-100, text);
}

// network spec "M0" (SURVEY.md App. B / beatrice_vst_b200/model_spec.py); the product's own
// statement of it -- the oracle and the torch cross-check each carry an independent one.
namespace spec {
struct Front {
  int k, cin, cout, stride;
};
constexpr Front kPhoneFront[6] = {{10, 1, 32, 5},   {3, 32, 64, 2},   {3, 64, 128, 2},
                                  {3, 128, 256, 2}, {3, 256, 256, 2}, {2, 256, 256, 2}};
constexpr int kPhoneDil[6] = {1, 2, 4, 1, 2, 4};
constexpr Front kPitchFront[6] = {{10, 1, 16, 5},  {3, 16, 32, 2},   {3, 32, 64, 2},
                                  {3, 64, 128, 2}, {3, 128, 128, 2}, {2, 128, 128, 2}};
constexpr int kPitchDil[3] = {1, 2, 4};
constexpr int kRates[4] = {5, 4, 4, 3};
constexpr int kStageCh[5] = {256, 128, 64, 32, 16};
constexpr int kMrfK[3] = {3, 7, 11};
constexpr int kMrfD[3] = {1, 3, 5};
constexpr int kPreK = 7, kPostK = 7;
}  // namespace spec

static size_t EncoderCount(const spec::Front* fs, int n_res, int width, int head_out) {
  size_t n = 0;
  for (int i = 0; i < 6; ++i) n += static_cast<size_t>(fs[i].k) * fs[i].cin * fs[i].cout + fs[i].cout;
  n += static_cast<size_t>(n_res) * (2 * width + 3 * width * width + width);
  n += static_cast<size_t>(width) * head_out + head_out;
  return n;
}
size_t PhoneParamCount(const FamilyDims& d) { return EncoderCount(spec::kPhoneFront, 6, 256, d.phone_channels); }
size_t PitchParamCount(const FamilyDims& d) {
  return EncoderCount(spec::kPitchFront, 3, 128, d.pitch_bins + kPitchFeatures);
}
size_t WaveParamCount(const FamilyDims& d) {
  size_t n = static_cast<size_t>(d.phone_channels) * kHidden + kHidden;
  n += static_cast<size_t>(d.pitch_bins) * kHidden + kPitchFeatures * kHidden;
  n += static_cast<size_t>(spec::kPreK) * kHidden * kHidden + kHidden;
  for (int s = 0; s < 4; ++s) {
    const int cin = spec::kStageCh[s], c = spec::kStageCh[s + 1];
    n += 2u * cin * spec::kRates[s] * c + c;
    for (int k : spec::kMrfK) n += 3u * 2u * (static_cast<size_t>(k) * c * c + c);
  }
  n += static_cast<size_t>(spec::kPostK) * 16 + 1;
  return n;
}
size_t SetterParamCount() {
  size_t n = 2u * (kHidden * kHidden + kHidden);
  for (int b = 0; b < kNBlocks; ++b) {
    const int c = spec::kStageCh[b + 1];
    n += kKvChannels + static_cast<size_t>(kKvChannels) * 2 * c + 2 * c;
  }
  return n;
}
size_t SpeakerPayloadFloats(const FamilyDims& d, uint32_t n) {
  if (d.has_setter)
    return static_cast<size_t>(kNFormant) * kHidden +
           static_cast<size_t>(n) * (static_cast<size_t>(kCodebookSize) * d.phone_channels + kHidden +
                                     static_cast<size_t>(kKvLength) * kKvChannels);
  return static_cast<size_t>(n) * kHidden;
}

int LoadFileBytes(const char* utf8_path, std::vector<uint8_t>* bytes) {
  FILE* f = std::fopen(utf8_path, "rb");
  if (!f) return 1;
  std::fseek(f, 0, SEEK_END);
  const long size = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (size < 0) {
    std::fclose(f);
    return 1;
  }
  bytes->resize(static_cast<size_t>(size));
  const size_t got = size > 0 ? std::fread(bytes->data(), 1, static_cast<size_t>(size), f) : 0;
  std::fclose(f);
  return got == static_cast<size_t>(size) ? 0 : 1;
}

int ParseFileImage(const void* data, size_t size, int family, uint32_t kind_a, uint32_t kind_b,
                   const std::function<long long(uint32_t)>& expected_floats, FileImage* out) {
  if (size < 16) return 2;
  uint32_t h[4];
  std::memcpy(h, data, 16);
  if (h[0] != kFileMagic || h[1] != static_cast<uint32_t>(family) || (h[2] != kind_a && h[2] != kind_b) ||
      (size - 16) % 4 != 0)
    return 4;
  const long long expect = expected_floats(h[3]);
  if (expect < 0) return 4;
  const long long have = static_cast<long long>((size - 16) / 4);
  if (have < expect) return 2;
  if (have > expect) return 3;
  out->family = h[1];
  out->kind = h[2];
  out->count = h[3];
  out->payload = reinterpret_cast<const float*>(static_cast<const uint8_t*>(data) + 16);
  out->n_floats = static_cast<size_t>(have);
  return 0;
}

// ---------------------------------------------------------------------------------------
// device plumbing
// ---------------------------------------------------------------------------------------
int UsableDeviceCount() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

int DefaultDevice() {
  static int dev = [] {
    const int n = UsableDeviceCount();
    if (n <= 0) return -1;
    const char* e = std::getenv("BEATRICE_B200_DEVICE");
    int d = e ? std::atoi(e) : 0;
    if (d < 0 || d >= n) return -1;   // an explicit, unusable choice is an error, not a silent move to device 0
    return d;
  }();
  // no device: latched like any other failure (silence out, error codes from the loaders) -- never a CPU fallback
  if (dev < 0) Fail(-101, "no usable CUDA device (or BEATRICE_B200_DEVICE out of range); this library has no CPU fallback", __FILE__, __LINE__);
  return dev;
}

namespace {
std::atomic<int> g_precision_override{-1};
}
void SetDefaultTcMode(int mode) { g_precision_override.store(mode >= 0 && mode <= 2 ? mode : -1, std::memory_order_relaxed); }

TcMode DefaultTcMode() {
  const int ov = g_precision_override.load(std::memory_order_relaxed);
  if (ov >= 0) return static_cast<TcMode>(ov);
  static TcMode mode = [] {
    const char* e = std::getenv("BEATRICE_B200_PRECISION");
    // default: split-bf16 on tcgen05 -- the fastest mode inside the <= 1e-4 RMS bar; f32 (CUDA cores) is opt-in
    if (!e) return kTcSplit;
    const std::string v(e);
    if (v == "bf16") return kTcBf16;
    if (v == "bf16x3" || v.empty()) return kTcSplit;
    if (v == "f32") return kTcOff;
    std::fprintf(stderr, "[libbeatrice_b200] warning: BEATRICE_B200_PRECISION=%s ignored (expected f32|bf16|bf16x3); using bf16x3\n", e);
    return kTcSplit;
  }();
  return mode;
}

void TcWeights::Pack(int device, const float* host_blob, const float* dev_blob, const std::vector<ConvW*>& convs) {
  std::vector<uint16_t> all;
  struct Slot {
    ConvW* c;
    size_t off_hi, off_lo;
  };
  std::vector<Slot> slots;
  for (ConvW* c : convs) {
    const bool ok = (c->cin == 16 || c->cin == 32 || c->cin == 64 || c->cin == 128 || c->cin == 256) && c->cout % 4 == 0;
    if (!ok) continue;
    int bn = 0, kc = 0;
    const size_t bytes = PackWeightsTc(nullptr, c->k, c->cin, c->cout, c->tc_bn_cap, &bn, &kc, nullptr, nullptr);
    const size_t n = bytes / sizeof(uint16_t);
    const size_t off = all.size();
    all.resize(off + 2 * n);
    PackWeightsTc(host_blob + (c->w - dev_blob), c->k, c->cin, c->cout, c->tc_bn_cap, &bn, &kc, all.data() + off,
                  all.data() + off + n);
    c->tc_bn = bn;
    c->tc_kc = kc;
    slots.push_back({c, off, off + n});
  }
  buf.Alloc(device, all.size() * sizeof(uint16_t), false);
  UploadSync(buf.p, all.data(), all.size() * sizeof(uint16_t));
  for (const Slot& sl : slots) {
    sl.c->tc_hi = buf.as<uint16_t>() + sl.off_hi;
    sl.c->tc_lo = buf.as<uint16_t>() + sl.off_lo;
  }
}

bool GraphsEnabled() {
  static bool on = [] {
    const char* e = std::getenv("BEATRICE_B200_NO_GRAPH");
    return !(e && e[0] == '1');
  }();
  return on;
}

bool ResStackEnabled() {
  static bool on = [] {
    const char* e = std::getenv("BEATRICE_B200_NO_FUSED_RESSTACK");
    return !(e && e[0] == '1');
  }();
  return on;
}

bool FusedMrfEnabled() {
  static bool on = [] {
    const char* e = std::getenv("BEATRICE_B200_NO_FUSED_MRF");
    return !(e && e[0] == '1');
  }();
  return on;
}

DeviceBuffer::~DeviceBuffer() { Free(); }
void DeviceBuffer::Free() {
  if (p) {
    int cur = 0;
    cudaGetDevice(&cur);
    if (device >= 0 && device != cur) cudaSetDevice(device);
    cudaFree(p);
    if (device >= 0 && device != cur) cudaSetDevice(cur);
  }
  p = nullptr;
  bytes = 0;
}
void DeviceBuffer::Alloc(int dev, size_t n_bytes, bool zero) {
  Free();
  device = dev;
  B200_CHECK(cudaSetDevice(dev));
  if (n_bytes == 0) n_bytes = 16;
  B200_CHECK(cudaMalloc(&p, n_bytes));
  bytes = n_bytes;
  if (zero) ZeroSync(p, n_bytes);
}

namespace {
struct Cursor {
  const float* host;
  const float* dev;
  size_t pos = 0;
  const float* Take(size_t n) {
    const float* r = dev + pos;
    pos += n;
    return r;
  }
  const float* HostAt(const float* dev_ptr) const { return host + (dev_ptr - dev); }
};
ConvW TakeConv(Cursor* c, int k, int cin, int cout) {
  ConvW w;
  w.k = k;
  w.cin = cin;
  w.cout = cout;
  w.w = c->Take(static_cast<size_t>(k) * cin * cout);
  w.b = c->Take(cout);
  return w;
}
void Upload(DeviceBuffer* buf, int dev, const float* host, size_t n) {
  buf->Alloc(dev, n * sizeof(float), false);
  UploadSync(buf->p, host, n * sizeof(float));
}
}  // namespace

// ---------------------------------------------------------------------------------------
// models
// ---------------------------------------------------------------------------------------
int EncoderModel::LoadFromImage(const void* data, size_t size, int on_device) {
  const size_t expect = is_pitch ? PitchParamCount(dims) : PhoneParamCount(dims);
  FileImage img;
  const int err = ParseFileImage(
      data, size, dims.family, is_pitch ? kKindPitch : kKindPhone, is_pitch ? kKindPitch : kKindPhone,
      [&](uint32_t cnt) { return cnt == expect ? static_cast<long long>(expect) : -1LL; }, &img);
  if (err) return err;
  device = on_device >= 0 ? on_device : DefaultDevice();
  Upload(&blob, device, img.payload, img.n_floats);
  Cursor c{img.payload, blob.as<float>()};
  const spec::Front* fs = is_pitch ? spec::kPitchFront : spec::kPhoneFront;
  const int* dl = is_pitch ? spec::kPitchDil : spec::kPhoneDil;
  n_res = is_pitch ? 3 : 6;
  width = is_pitch ? 128 : 256;
  head_out = is_pitch ? dims.pitch_bins + kPitchFeatures : dims.phone_channels;
  for (int i = 0; i < 6; ++i) {
    front[i] = TakeConv(&c, fs[i].k, fs[i].cin, fs[i].cout);
    stride[i] = fs[i].stride;
  }
  for (int i = 0; i < n_res; ++i) {
    gamma[i] = c.Take(width);
    beta[i] = c.Take(width);
    res[i] = TakeConv(&c, 3, width, width);
    dil[i] = dl[i];
  }
  head = TakeConv(&c, 1, width, head_out);
  {
    std::vector<ConvW*> convs;
    for (int i = 1; i < 6; ++i) convs.push_back(&front[i]);
    for (int i = 0; i < n_res; ++i) convs.push_back(&res[i]);
    convs.push_back(&head);
    for (ConvW* cv : convs) cv->tc_bn_cap = 64;   // few rows per hop (T <= 16): favour CTA count
    // front-end layers 1-2 (16 / 8 rows per stream, N = 64 / 128 at the content encoder): 32-column tiles -- twice the
    // CTAs, half the epilogue (bias + GELU + bf16 split) per thread: 15 -> 13 us each
    front[1].tc_bn_cap = front[2].tc_bn_cap = 32;
    if (const char* ev = std::getenv("BEATRICE_B200_FE_BN")) {   // developer overrides: layers 1-2 / layers 3-4
      const int cap = std::atoi(ev);
      if (cap == 64 || cap == 32 || cap == 16) front[1].tc_bn_cap = front[2].tc_bn_cap = cap;
    }
    if (const char* ev = std::getenv("BEATRICE_B200_FE_BN34")) {
      const int cap = std::atoi(ev);
      if (cap == 64 || cap == 32) front[3].tc_bn_cap = front[4].tc_bn_cap = cap;
    }
    tc.Pack(device, img.payload, blob.as<float>(), convs);
  }
  rs_ok = ResStackSupported(width, n_res, dil) && front[5].k == 2 && stride[5] == 2 && front[5].cin == width && front[5].cout == width;
  if (rs_ok) {   // fused encoder chain (b200_enc.cu): front layer 5, the residual blocks, and the head where it fits
    rs_head = !is_pitch && head.k == 1 && (head_out == 128 || head_out == 256) && head_out <= width;
    std::vector<ChainLayer> layers;
    rs_n_blk = 0;
    auto add = [&](int kind, const ConvW& cw, int d) {
      layers.push_back({c.HostAt(cw.w), cw.k, cw.cout});
      rs_kind[rs_n_blk] = kind;
      rs_dil[rs_n_blk] = d;
      ++rs_n_blk;
    };
    add(1, front[5], 1);
    for (int i = 0; i < n_res; ++i) add(0, res[i], dil[i]);
    if (rs_head) add(2, head, 1);
    std::vector<float> par(static_cast<size_t>(3) * rs_n_blk * width, 0.0f);
    auto put = [&](int which, int blk, const float* dev_ptr, int n) {
      std::memcpy(par.data() + (static_cast<size_t>(which) * rs_n_blk + blk) * width, c.HostAt(dev_ptr), sizeof(float) * n);
    };
    put(0, 0, front[5].b, width);
    for (int i = 0; i < n_res; ++i) {
      put(0, 1 + i, res[i].b, width);
      put(1, 1 + i, gamma[i], width);
      put(2, 1 + i, beta[i], width);
    }
    if (rs_head) put(0, 1 + n_res, head.b, head_out);
    std::vector<uint16_t> packed(PackChainWeights(layers.data(), rs_n_blk, width, nullptr));
    PackChainWeights(layers.data(), rs_n_blk, width, packed.data());
    rs_w.Alloc(device, packed.size() * sizeof(uint16_t), false);
    UploadSync(rs_w.p, packed.data(), packed.size() * sizeof(uint16_t));
    Upload(&rs_par, device, par.data(), par.size());
    rs_bias = rs_par.as<float>();
    rs_gamma = rs_bias + static_cast<size_t>(rs_n_blk) * width;
    rs_beta = rs_gamma + static_cast<size_t>(rs_n_blk) * width;
  }
  ++generation;
  loaded = true;
  return 0;
}
int EncoderModel::LoadFromFile(const char* path, int on_device) {
  std::vector<uint8_t> bytes;
  if (const int e = LoadFileBytes(path, &bytes)) return e;
  return LoadFromImage(bytes.data(), bytes.size(), on_device);
}

int WaveModel::LoadFromImage(const void* data, size_t size, int on_device) {
  const size_t expect = WaveParamCount(dims);
  FileImage img;
  const int err = ParseFileImage(data, size, dims.family, kKindWavegen, kKindWavegen,
                                 [&](uint32_t cnt) { return cnt == expect ? static_cast<long long>(expect) : -1LL; },
                                 &img);
  if (err) return err;
  device = on_device >= 0 ? on_device : DefaultDevice();
  Upload(&blob, device, img.payload, img.n_floats);
  Cursor c{img.payload, blob.as<float>()};
  embed = TakeConv(&c, 1, dims.phone_channels, kHidden);
  pitch_emb = c.Take(static_cast<size_t>(dims.pitch_bins) * kHidden);
  feat_proj = c.Take(kPitchFeatures * kHidden);
  pre = TakeConv(&c, spec::kPreK, kHidden, kHidden);
  // ConvTranspose1d bias [C_out] is replicated per output phase for the 2-tap GEMM form
  std::vector<float> rep;
  size_t rep_off[4];
  for (int s = 0; s < 4; ++s) {
    const int cin = spec::kStageCh[s], co = spec::kStageCh[s + 1], r = spec::kRates[s];
    ConvW u;
    u.k = 2;
    u.cin = cin;
    u.cout = r * co;
    u.w = c.Take(2u * cin * r * co);
    const float* b_dev = c.Take(co);
    const float* b_host = c.HostAt(b_dev);
    ups_bias_raw[s] = b_dev;
    rep_off[s] = rep.size();
    for (int p = 0; p < r; ++p) rep.insert(rep.end(), b_host, b_host + co);
    ups[s] = u;
    for (int ki = 0; ki < 3; ++ki)
      for (int di = 0; di < 3; ++di) {
        c1[s][ki][di] = TakeConv(&c, spec::kMrfK[ki], co, co);
        c2[s][ki][di] = TakeConv(&c, spec::kMrfK[ki], co, co);
      }
  }
  post = TakeConv(&c, spec::kPostK, 16, 1);
  Upload(&ups_bias, device, rep.data(), rep.size());
  {
    std::vector<ConvW*> convs;
    pre.tc_bn_cap = 64;
    ups[0].tc_bn_cap = 128;
    convs.push_back(&pre);
    for (int s = 0; s < 4; ++s) {
      convs.push_back(&ups[s]);
      for (int ki = 0; ki < 3; ++ki)
        for (int di = 0; di < 3; ++di) {
          convs.push_back(&c1[s][ki][di]);
          convs.push_back(&c2[s][ki][di]);
        }
    }
    tc.Pack(device, img.payload, blob.as<float>(), convs);
  }
  for (int s = 0; s < 4; ++s) ups[s].b = ups_bias.as<float>() + rep_off[s];
  {
    const int d3 = 3;   // the chain kernel sizes a k = 7 block's history like a residual block of dilation 3 (6 rows)
    pre_chain_ok = pre.k == 7 && pre.cin == kHidden && pre.cout == kHidden && ResStackSupported(kHidden, 1, &d3);
    if (pre_chain_ok) {
      const ChainLayer layer = {c.HostAt(pre.w), 7, kHidden};
      std::vector<uint16_t> packed(PackChainWeights(&layer, 1, kHidden, nullptr));
      PackChainWeights(&layer, 1, kHidden, packed.data());
      pre_w.Alloc(device, packed.size() * sizeof(uint16_t), false);
      UploadSync(pre_w.p, packed.data(), packed.size() * sizeof(uint16_t));
      std::vector<float> par(static_cast<size_t>(3) * kHidden, 0.0f);   // bias | gamma (unused) | beta (unused)
      std::memcpy(par.data(), c.HostAt(pre.b), sizeof(float) * kHidden);
      Upload(&pre_par, device, par.data(), par.size());
      // two-block chain: conditioning (phone embedding, rows past the phone channels zero) + pre conv
      cond_chain_ok = embed.k == 1 && embed.cout == kHidden && embed.cin <= kHidden && embed.cin % 4 == 0;
      if (cond_chain_ok) {
        std::vector<float> we(static_cast<size_t>(kHidden) * kHidden, 0.0f);
        std::memcpy(we.data(), c.HostAt(embed.w), sizeof(float) * embed.cin * kHidden);
        const ChainLayer layers[2] = {{we.data(), 1, kHidden}, {c.HostAt(pre.w), 7, kHidden}};
        std::vector<uint16_t> packed2(PackChainWeights(layers, 2, kHidden, nullptr));
        PackChainWeights(layers, 2, kHidden, packed2.data());
        cond_pre_w.Alloc(device, packed2.size() * sizeof(uint16_t), false);
        UploadSync(cond_pre_w.p, packed2.data(), packed2.size() * sizeof(uint16_t));
        std::vector<float> par2(static_cast<size_t>(3) * 2 * kHidden, 0.0f);   // bias[2][C] | gamma[2][C] | beta[2][C]
        std::memcpy(par2.data(), c.HostAt(embed.b), sizeof(float) * kHidden);
        std::memcpy(par2.data() + kHidden, c.HostAt(pre.b), sizeof(float) * kHidden);
        Upload(&cond_pre_par, device, par2.data(), par2.size());
      }
    }
  }
  {
    // fused MRF kernel images: for every stage the kernel has a form for, both precisions
    std::vector<float> bias_all;
    size_t bias_off[4][3] = {};
    std::vector<uint16_t> packed[2];
    size_t w_off[2][4][3] = {};
    bool have[4] = {};
    for (int s = 0; s < 4; ++s) {
      const int co = spec::kStageCh[s + 1];
      if (co != 16 && co != 32 && co != 64 && co != 128) continue;
      have[s] = true;
      for (int ki = 0; ki < 3; ++ki) {
        const float* wsrc[6];
        bias_off[s][ki] = bias_all.size();
        for (int di = 0; di < 3; ++di) {
          wsrc[2 * di] = c.HostAt(c1[s][ki][di].w);
          wsrc[2 * di + 1] = c.HostAt(c2[s][ki][di].w);
          const float* b1 = c.HostAt(c1[s][ki][di].b);
          const float* b2 = c.HostAt(c2[s][ki][di].b);
          bias_all.insert(bias_all.end(), b1, b1 + co);
          bias_all.insert(bias_all.end(), b2, b2 + co);
        }
        for (int sp = 0; sp < 2; ++sp) {
          // rows [W_hi ; W_lo] (2-MMA split scheme) wherever the single-CTA kernel runs the branch: every
          // branch at C <= 64, and the k = 3 branch of stage 0 (C = 128), which runs beside the clusters
          const bool concat = co <= 64 || ki == 0;
          const size_t n = PackMrfWeights(wsrc, spec::kMrfK[ki], co, sp == 1, concat, nullptr);
          size_t off = (packed[sp].size() + 127) / 128 * 128;   // 256-byte aligned images
          packed[sp].resize(off + n);
          PackMrfWeights(wsrc, spec::kMrfK[ki], co, sp == 1, concat, packed[sp].data() + off);
          w_off[sp][s][ki] = off;
        }
      }
    }
    {
      // upsamplers of stages 1..3 as images for the fused kernel's prologue
      std::vector<uint16_t> img;
      size_t off[4] = {};
      for (int s = 1; s < 4; ++s) {
        const int co = spec::kStageCh[s + 1], r = spec::kRates[s];
        ups_img_ptr[s] = nullptr;
        if (!have[s] || spec::kStageCh[s] != 2 * co) continue;
        off[s] = (img.size() + 127) / 128 * 128;
        img.resize(off[s] + PackMrfUpsWeights(c.HostAt(ups[s].w), co, r, nullptr));
        PackMrfUpsWeights(c.HostAt(ups[s].w), co, r, img.data() + off[s]);
      }
      if (!img.empty()) {
        ups_img.Alloc(device, img.size() * sizeof(uint16_t), false);
        UploadSync(ups_img.p, img.data(), img.size() * sizeof(uint16_t));
        for (int s = 1; s < 4; ++s)
          if (have[s] && spec::kStageCh[s] == 2 * spec::kStageCh[s + 1]) ups_img_ptr[s] = ups_img.as<uint16_t>() + off[s];
      }
    }
    Upload(&mrf_bias, device, bias_all.data(), bias_all.size());
    for (int sp = 0; sp < 2; ++sp) {
      mrf_w[sp].Alloc(device, packed[sp].size() * sizeof(uint16_t), false);
      if (!packed[sp].empty())
        UploadSync(mrf_w[sp].p, packed[sp].data(), packed[sp].size() * sizeof(uint16_t));
    }
    for (int s = 0; s < 4; ++s)
      for (int ki = 0; ki < 3; ++ki) {
        mrf_bias_ptr[s][ki] = have[s] ? mrf_bias.as<float>() + bias_off[s][ki] : nullptr;
        for (int sp = 0; sp < 2; ++sp) mrf_w_ptr[sp][s][ki] = have[s] ? mrf_w[sp].as<uint16_t>() + w_off[sp][s][ki] : nullptr;
      }
  }
  ++generation;
  loaded = true;
  return 0;
}
int WaveModel::LoadFromFile(const char* path, int on_device) {
  std::vector<uint8_t> bytes;
  if (const int e = LoadFileBytes(path, &bytes)) return e;
  return LoadFromImage(bytes.data(), bytes.size(), on_device);
}

int SetterModel::LoadFromImage(const void* data, size_t size, int on_device) {
  const size_t expect = SetterParamCount();
  FileImage img;
  const int err = ParseFileImage(data, size, dims.family, kKindSetter, kKindSetter,
                                 [&](uint32_t cnt) { return cnt == expect ? static_cast<long long>(expect) : -1LL; },
                                 &img);
  if (err) return err;
  device = on_device >= 0 ? on_device : DefaultDevice();
  Upload(&blob, device, img.payload, img.n_floats);
  Cursor c{img.payload, blob.as<float>()};
  add_w = c.Take(kHidden * kHidden);
  add_b = c.Take(kHidden);
  for_w = c.Take(kHidden * kHidden);
  for_b = c.Take(kHidden);
  for (int b = 0; b < kNBlocks; ++b) {
    const int ch = spec::kStageCh[b + 1];
    query[b] = c.Take(kKvChannels);
    film_w[b] = c.Take(static_cast<size_t>(kKvChannels) * 2 * ch);
    film_b[b] = c.Take(2 * ch);
  }
  loaded = true;
  return 0;
}
int SetterModel::LoadFromFile(const char* path, int on_device) {
  std::vector<uint8_t> bytes;
  if (const int e = LoadFileBytes(path, &bytes)) return e;
  return LoadFromImage(bytes.data(), bytes.size(), on_device);
}

// ---------------------------------------------------------------------------------------
// state arena
// ---------------------------------------------------------------------------------------
static int SlotsFor(int history_rows, int T) {
  const int slots = history_rows > 0 ? (history_rows + T - 1) / T + 1 : 1;
  if (slots > 16) Fail(-102, "ring needs more than 16 slots", __FILE__, __LINE__);
  return slots;
}
int StateArena::Plan(int history_rows, int T, int C) {
  Ring r;
  r.T = T;
  r.C = C;
  r.slots = SlotsFor(history_rows, T);
  rings_.push_back(r);
  return static_cast<int>(rings_.size()) - 1;
}
int StateArena::PlanH(int history_rows, int T, int C, bool with_lo) {
  Ring r;
  r.T = T;
  r.C = C;
  r.slots = SlotsFor(history_rows, T);
  r.is_bf16 = true;
  r.has_lo = with_lo;
  rings_.push_back(r);
  return static_cast<int>(rings_.size()) - 1;
}
void StateArena::Commit(int device, int B) {
  B_ = B;
  size_t total = 0;
  offsets_.resize(rings_.size());
  for (size_t i = 0; i < rings_.size(); ++i) {
    offsets_[i] = total;
    const size_t elems = rings_[i].StreamStride() * B;
    size_t bytes = rings_[i].is_bf16 ? elems * 2 * (rings_[i].has_lo ? 2 : 1) : elems * 4;
    bytes = (bytes + 255) / 256 * 256;  // keep every ring 256-byte aligned
    total += bytes;
  }
  buf_.Alloc(device, total, true);
  frame_.Alloc(device, sizeof(int), true);
  for (size_t i = 0; i < rings_.size(); ++i) {
    uint8_t* p = buf_.as<uint8_t>() + offsets_[i];
    Ring& r = rings_[i];
    if (r.is_bf16) {
      r.hi = reinterpret_cast<uint16_t*>(p);
      if (r.has_lo) r.lo = r.hi + r.StreamStride() * B;
    } else {
      r.base = reinterpret_cast<float*>(p);
    }
  }
}
void StateArena::Clear() {
  rings_.clear();
  offsets_.clear();
  buf_.Free();
  frame_.Free();
  B_ = 0;
}
void StateArena::ZeroAll(cudaStream_t s) { B200_CHECK(cudaMemsetAsync(buf_.p, 0, buf_.bytes, s)); }
void StateArena::ZeroStream(int b, cudaStream_t s) {
  for (const Ring& r : rings_) {
    const size_t n = r.StreamStride();
    if (r.is_bf16) {
      B200_CHECK(cudaMemsetAsync(r.hi + n * b, 0, n * 2, s));
      if (r.lo) B200_CHECK(cudaMemsetAsync(r.lo + n * b, 0, n * 2, s));
    } else {
      B200_CHECK(cudaMemsetAsync(r.base + n * b, 0, n * 4, s));
    }
  }
}

// ---------------------------------------------------------------------------------------
// program builders
// ---------------------------------------------------------------------------------------
namespace {

struct DescBuilder {
  std::vector<ConvDesc> host;
  int Add(const ConvDesc& d) {
    host.push_back(d);
    return static_cast<int>(host.size()) - 1;
  }
};

ConvDesc MakeConv(const Ring& x, const ConvW& w, int dil, int stride, int T_out, const Ring& y, int in_act,
                  int out_act) {
  ConvDesc d;
  std::memset(&d, 0, sizeof(d));
  d.x[0] = x.base;
  d.x[1] = d.x[2] = nullptr;
  d.n_x = 1;
  d.in_scale = 1.0f;
  d.x_slots = x.slots;
  d.x_T = x.T;
  d.x_C = x.C;
  d.in_act = in_act;
  d.w = w.w;
  d.bias = w.b;
  d.k = w.k;
  d.dil = dil;
  d.stride = stride;
  d.C_in = w.cin;
  d.N = w.cout;
  d.T = T_out;
  d.y = y.base;
  d.y_slots = y.slots;
  d.y_T = y.T;
  d.y_C = y.C;
  d.out_act = out_act;
  d.w_tc = w.tc_hi;
  d.w_tc_lo = w.tc_lo;
  d.tc_bn = w.tc_bn;
  d.tc_kc = w.tc_kc;
  return d;
}

// input taken from a bf16 ring by cp.async (tensor-core path only)
void SetInH(ConvDesc* d, const Ring& h) {
  d->xh = h.hi;
  d->xl = h.lo;
  d->x_slots = h.slots;
  d->x_T = h.T;
  d->x_C = h.C;
}
// the epilogue additionally stores bf16(act(v)) for the tensor-core consumer
void SetOutH(ConvDesc* d, const Ring& h, int act) {
  d->yh = h.hi;
  d->yl = h.lo;
  d->yh_slots = h.slots;
  d->yh_act = act;
}

void SetRes(ConvDesc* d, const Ring& r) {
  d->res = r.base;
  d->res_slots = r.slots;
  d->res_T = r.T;
}

double ConvFlops(const ConvDesc& d, int B) { return 2.0 * d.C_in * d.k * d.N * d.T * B; }
double ConvBytes(const ConvDesc& d, int B) {
  const double w = 4.0 * d.k * d.C_in * d.N;
  const double in = 4.0 * B * d.T * d.stride * d.C_in * d.n_x;
  const double out = 4.0 * B * d.T * d.N * (d.res ? 2 : 1);
  return w + in + out;
}

// CUDA-core or tensor-core launcher for one (possibly z-batched) conv GEMM.
std::function<void(cudaStream_t)> GemmLauncher(const ConvDesc* dp, const ConvDesc& h, int nz, int B, const int* frame,
                                               TcMode tc, bool tc_allowed) {
  if (tc != kTcOff && tc_allowed && h.w_tc != nullptr) {
    const bool split = tc == kTcSplit;
    return [=](cudaStream_t s) { LaunchConvGemmTc(dp, h, nz, B, frame, split, s); };
  }
  return [=](cudaStream_t s) { LaunchConvGemm(dp, h, nz, B, frame, s); };
}

// developer aid: BEATRICE_B200_TC_TRACE=<substring of an op name> makes that op's kernel print a timeline
ConvDesc Traced(ConvDesc h, const std::string& op_name) {
  const char* e = std::getenv("BEATRICE_B200_TC_TRACE");
  h.trace = (e && e[0] && op_name.find(e) != std::string::npos) ? 1 : 0;
  return h;
}

Ring FlatRing(float* base, int T, int C) {
  Ring r;
  r.base = base;
  r.slots = 1;
  r.T = T;
  r.C = C;
  return r;
}

}  // namespace

void EncoderState::Build(const EncoderModel* m, int B_, int device_, const float* external_stage, TcMode tc) {
  B = B_;
  device = device_;
  model = m;
  model_generation = m->generation;
  program.clear();
  arena.Clear();
  B200_CHECK(cudaSetDevice(device));
  // the encoders feed the discrete pitch arg-max: whenever tensor cores are on they run the
  // near-fp32 split-bf16 form, also in "bf16" mode (plain bf16 is for the vocoder only)
  const bool tcm = tc != kTcOff;
  if (tcm) tc = kTcSplit;

  int t_in[6], t_out[6];
  int t = kInHop;
  for (int i = 0; i < 6; ++i) {
    t_in[i] = t;
    t /= m->stride[i];
    t_out[i] = t;
  }
  // ring_in[i] = input of front-end layer i.  Tensor-core mode: layers 1..5 read bf16 rings.
  int ring_in[6];
  for (int i = 0; i < 6; ++i) {
    const int hist = m->front[i].k - m->stride[i];
    ring_in[i] = (tcm && i >= 1) ? arena.PlanH(hist, t_in[i], m->front[i].cin, true)
                                 : arena.Plan(hist, t_in[i], m->front[i].cin);
  }
  // tensor-core mode: the residual stack (ChanNorm + GELU + dilated conv, n_res blocks) is ONE cluster kernel
  // (b200_enc.cu); it keeps its own conv-input histories, so only the stack's input and output rows are rings
  const bool rs_fused = tcm && m->rs_ok && ResStackEnabled();
  std::vector<int> ring_x(m->n_res + 1, -1), ring_g(m->n_res, -1);
  ring_x[0] = arena.Plan(0, 1, m->width);
  for (int r = 0; r < m->n_res; ++r) {
    if (!rs_fused) ring_g[r] = tcm ? arena.PlanH(2 * m->dil[r], 1, m->width, true) : arena.Plan(2 * m->dil[r], 1, m->width);
    if (!rs_fused || r == m->n_res - 1) ring_x[r + 1] = arena.Plan(0, 1, m->width);
  }
  const int ring_xh = tcm ? arena.PlanH(0, 1, m->width, true) : -1;   // bf16 copy of the last x for the head
  arena.Commit(device, B);
  if (external_stage) {
    in_stage.Free();
    stage_ptr = external_stage;
  } else {
    in_stage.Alloc(device, sizeof(float) * B * kInHop, true);
    stage_ptr = in_stage.as<float>();
  }
  head_out.Alloc(device, sizeof(float) * B * m->head_out, true);

  DescBuilder db;
  std::vector<int> conv_idx;
  for (int i = 0; i < 6; ++i) {
    const Ring& x = arena.ring(ring_in[i]);
    const Ring& y = (i < 5) ? arena.ring(ring_in[i + 1]) : arena.ring(ring_x[0]);
    ConvDesc d = MakeConv(x, m->front[i], 1, m->stride[i], t_out[i], y, kActNone, kActGelu);
    if (x.is_bf16) SetInH(&d, x);
    if (y.is_bf16) SetOutH(&d, y, kActNone);   // v is already GELU'd by out_act
    conv_idx.push_back(db.Add(d));
  }
  std::vector<int> res_idx;
  for (int r = 0; r < m->n_res && !rs_fused; ++r) {
    const Ring& g = arena.ring(ring_g[r]);
    ConvDesc d = MakeConv(g, m->res[r], m->dil[r], 1, 1, arena.ring(ring_x[r + 1]), kActNone, kActNone);
    if (g.is_bf16) SetInH(&d, g);
    SetRes(&d, arena.ring(ring_x[r]));
    if (ring_xh >= 0 && r == m->n_res - 1) SetOutH(&d, arena.ring(ring_xh), kActNone);
    res_idx.push_back(db.Add(d));
  }
  const Ring head_ring = FlatRing(head_out.as<float>(), 1, m->head_out);
  ConvDesc hd = MakeConv(arena.ring(ring_x[m->n_res]), m->head, 1, 1, 1, head_ring, kActNone, kActNone);
  if (ring_xh >= 0) SetInH(&hd, arena.ring(ring_xh));
  const int head_idx = db.Add(hd);

  descs.Alloc(device, sizeof(ConvDesc) * db.host.size(), false);
  UploadSync(descs.p, db.host.data(), sizeof(ConvDesc) * db.host.size());
  const ConvDesc* dd = descs.as<ConvDesc>();
  const int* frame = arena.frame();
  const int Bn = B;
  const char* tag = m->is_pitch ? "pitch" : "phone";

  const bool fe0_fused = Frontend0Supported(db.host[conv_idx[0]]);
  if (!fe0_fused) {
    const Ring in0 = arena.ring(ring_in[0]);
    const float* stage = stage_ptr;
    Op op;
    op.name = std::string(tag) + ".ingest";
    op.bytes = 8.0 * B * kInHop;
    op.launch = [=](cudaStream_t s) { LaunchIngest(stage, in0.base, in0.slots, in0.T, in0.C, Bn, frame, s); };
    program.push_back(op);
  }
  for (int i = 0; i < 6; ++i) {
    const ConvDesc h = Traced(db.host[conv_idx[i]], std::string(tag) + ".fe" + std::to_string(i));
    const ConvDesc* dp = dd + conv_idx[i];
    Op op;
    op.name = std::string(tag) + ".fe" + std::to_string(i);
    op.flops = ConvFlops(h, B);
    op.bytes = ConvBytes(h, B);
    if (i == 0 && fe0_fused) {
      const float* stage = stage_ptr;
      float* ring0 = arena.ring(ring_in[0]).base;
      op.launch = [=](cudaStream_t s) { LaunchFrontend0(h, stage, ring0, Bn, frame, s); };   // ingest + conv
    } else if (i == 0)
      op.launch = [=](cudaStream_t s) { LaunchDirectConv(dp, h, Bn, frame, s); };
    else
      op.launch = GemmLauncher(dp, h, 1, Bn, frame, tc, tcm);
    if (!(rs_fused && i == 5)) program.push_back(op);   // layer 5 is the first block of the fused chain
  }
  n_rs_blocks = 0;
  if (rs_fused) {
    const size_t n_hist = ResStackHistElems(m->width, m->n_res, m->dil, B);
    rs_hist.Alloc(device, n_hist * sizeof(uint16_t), true);
    std::vector<MrfHistBlock> blocks;
    ResStackHistBlocks(m->width, m->n_res, m->dil, B, rs_hist.as<uint16_t>(), &blocks);
    n_rs_blocks = static_cast<int>(blocks.size());
    rs_blocks.Alloc(device, sizeof(MrfHistBlock) * blocks.size(), false);
    UploadSync(rs_blocks.p, blocks.data(), sizeof(MrfHistBlock) * blocks.size());
    ResStackParams rp;
    std::memset(&rp, 0, sizeof(rp));
    const Ring& fin = arena.ring(ring_in[5]);           // front layer 5 reads the hop's two rows of layer 4's output
    rp.fin_h = fin.hi;
    rp.fin_l = fin.lo;
    if (!m->rs_head) {                                  // a separate head conv reads the stack's output
      rp.x_out = arena.ring(ring_x[m->n_res]).base;
      rp.xh_out = arena.ring(ring_xh).hi;
      rp.xl_out = arena.ring(ring_xh).lo;
    } else {
      rp.head_out = head_out.as<float>();
      rp.head_out2 = head_copy;
      head_dual = head_copy != nullptr;
      rp.head_n = m->head_out;
    }
    rp.w = m->rs_w.as<uint16_t>();
    rp.bias = m->rs_bias;
    rp.gamma = m->rs_gamma;
    rp.beta = m->rs_beta;
    rp.hist = rs_hist.as<uint16_t>();
    rp.n_blk = m->rs_n_blk;
    for (int r = 0; r < m->rs_n_blk; ++r) {
      rp.kind[r] = m->rs_kind[r];
      rp.dil[r] = m->rs_dil[r];
    }
    rp.B = B;
    rp.n_tiles = ResStackTiles(B);
    if (const char* ev = std::getenv("BEATRICE_B200_ENC_TRACE")) rp.trace = std::atoi(ev);
    const int width = m->width;
    Op op;
    op.name = std::string(tag) + ".chain";   // fe5 + residual stack (+ head)
    op.flops = 2.0 * 3 * width * width * B * m->n_res + 2.0 * 2 * width * width * B + (m->rs_head ? 2.0 * width * m->head_out * B : 0.0);
    op.bytes = 4.0 * 3 * width * width * m->n_res + 8.0 * B * width;
    op.launch = [=](cudaStream_t s) { LaunchResStack(rp, width, s); };
    program.push_back(op);
  }
  for (int r = 0; r < m->n_res && !rs_fused; ++r) {
    NormDesc nd;
    std::memset(&nd, 0, sizeof(nd));
    const Ring& xr = arena.ring(ring_x[r]);
    const Ring& gr = arena.ring(ring_g[r]);
    nd.x = xr.base;
    nd.x_slots = xr.slots;
    nd.T = 1;
    nd.C = m->width;
    nd.gamma = m->gamma[r];
    nd.beta = m->beta[r];
    nd.y = gr.base;      // nullptr in tensor-core mode: only the bf16 planes are kept
    nd.y_slots = gr.slots;
    nd.yh = gr.hi;
    nd.yl = gr.lo;
    nd.yh_slots = gr.slots;
    Op on;
    on.name = std::string(tag) + ".res" + std::to_string(r) + ".norm";
    on.bytes = 8.0 * B * m->width;
    on.launch = [=](cudaStream_t s) { LaunchNorm(nd, Bn, frame, s); };
    program.push_back(on);
    const ConvDesc h = Traced(db.host[res_idx[r]], std::string(tag) + ".res" + std::to_string(r) + ".conv");
    const ConvDesc* dp = dd + res_idx[r];
    Op oc;
    oc.name = std::string(tag) + ".res" + std::to_string(r) + ".conv";
    oc.flops = ConvFlops(h, B);
    oc.bytes = ConvBytes(h, B);
    oc.launch = GemmLauncher(dp, h, 1, Bn, frame, tc, tcm);
    program.push_back(oc);
  }
  {
    const ConvDesc h = Traced(db.host[head_idx], std::string(tag) + ".head");
    const ConvDesc* dp = dd + head_idx;
    Op op;
    op.name = std::string(tag) + ".head";
    op.flops = ConvFlops(h, B);
    op.bytes = ConvBytes(h, B);
    op.launch = GemmLauncher(dp, h, 1, Bn, frame, tc, tcm);
    if (!(rs_fused && m->rs_head)) program.push_back(op);   // the head is the last block of the fused chain
  }
  {
    int* f = arena.frame();
    Op op;
    op.name = std::string(tag) + ".advance";
    op.launch = [=](cudaStream_t s) { LaunchAdvance(f, s); };
    program.push_back(op);
  }
}

void EncoderState::ZeroStream(int b, cudaStream_t s) {
  arena.ZeroStream(b, s);
  LaunchMrfZeroStream(rs_blocks.as<MrfHistBlock>(), n_rs_blocks, b, s);
}
void EncoderState::ZeroAll(cudaStream_t s) {
  arena.ZeroAll(s);
  if (n_rs_blocks > 0 && rs_hist.p) B200_CHECK(cudaMemsetAsync(rs_hist.p, 0, rs_hist.bytes, s));
}

void WaveState::AllocCond(const FamilyDims& dims, int B_, int device_) {
  if (cond_ready && B == B_ && device == device_) return;
  B = B_;
  device = device_;
  phone_in.Alloc(device, sizeof(float) * B * dims.phone_channels, true);
  q_in.Alloc(device, sizeof(int) * B, true);
  feat_in.Alloc(device, sizeof(float) * B * kPitchFeatures, true);
  spk.Alloc(device, sizeof(float) * B * kHidden, true);
  formant.Alloc(device, sizeof(float) * B * kHidden, true);
  for (int s = 0; s < 4; ++s) film[s].Alloc(device, sizeof(float) * B * 2 * spec::kStageCh[s + 1], true);
  out.Alloc(device, sizeof(float) * B * kOutHop, true);
  cond_ready = true;
}

namespace {

// How the three MRF branches of a stage map to launches and CTAs of the single-CTA kernel (b200_mrf.cu).
// Text form: launches separated by '/', CTA classes (blockIdx.y) of a launch by '|', the branches ONE CTA runs one
// after the other by ',', each branch named by its kernel size; "@KB" after a launch = request at least that much
// dynamic shared memory (fewer CTAs per SM / none of another launch beside them).  Every branch exactly once.
//   "11|7|3"        one launch, three CTA classes (round 1)
//   "11|7,3"        one launch, two classes of 11 and 10 taps per conv
//   "11@128/7|3"    a PDL pair: k = 11 alone on its SMs, k = 7 and k = 3 sharing the others
struct MrfLaunchPlan {
  int n_launch = 0;
  struct L {
    int n_y = 0, ylen[3] = {0, 0, 0}, yseq[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, smem_kb = 0;
  } l[3];
};

bool ParseMrfPlan(const std::string& text, MrfLaunchPlan* out) {
  MrfLaunchPlan pl;
  int seen = 0;
  size_t i = 0;
  pl.n_launch = 1;
  pl.l[0].n_y = 1;
  auto cur = [&]() -> MrfLaunchPlan::L& { return pl.l[pl.n_launch - 1]; };
  while (i < text.size()) {
    const char ch = text[i];
    if (ch >= '0' && ch <= '9') {
      int v = 0;
      while (i < text.size() && text[i] >= '0' && text[i] <= '9') v = v * 10 + (text[i++] - '0');
      const int ki = v == 3 ? 0 : (v == 7 ? 1 : (v == 11 ? 2 : -1));
      MrfLaunchPlan::L& L = cur();
      if (ki < 0 || (seen >> ki) & 1 || L.ylen[L.n_y - 1] >= 3) return false;
      seen |= 1 << ki;
      L.yseq[L.n_y - 1][L.ylen[L.n_y - 1]++] = ki;
      continue;
    }
    if (ch == ',') {
      ++i;
    } else if (ch == '|') {
      if (cur().ylen[cur().n_y - 1] == 0 || cur().n_y >= 3) return false;
      ++cur().n_y;
      ++i;
    } else if (ch == '@') {
      int v = 0;
      ++i;
      while (i < text.size() && text[i] >= '0' && text[i] <= '9') v = v * 10 + (text[i++] - '0');
      cur().smem_kb = v;
    } else if (ch == '/') {
      if (cur().ylen[cur().n_y - 1] == 0 || pl.n_launch >= 3) return false;
      ++pl.n_launch;
      cur().n_y = 1;
      ++i;
    } else {
      return false;
    }
  }
  if (seen != 7 || cur().ylen[cur().n_y - 1] == 0) return false;
  *out = pl;
  return true;
}

// default plans (measured on B200 at 256 streams) and the developer override
// BEATRICE_B200_MRF_PLAN="<C=64 plan>;<C=32 plan>;<C=16 plan>" (an empty field keeps the default)
MrfLaunchPlan MrfPlanFor(int c) {
  // (measured at 256 streams: "11|7,3" costs 2-10 us per stage -- a chain of twelve convs is longer than the k = 11 chain of
  //  six, per-conv hand-off latency and not the tap count bounds these CTAs -- and "11@128/7|3" 11 us at stage 2: the k = 7 / k = 3
  //  CTAs then queue for the SMs the k = 11 CTAs leave; the three-class single launch stays the default everywhere)
  std::string text = "11|7|3";
  (void)c;
  if (const char* ev = std::getenv("BEATRICE_B200_MRF_PLAN")) {
    std::string all(ev), field;
    const int want = c >= 64 ? 0 : (c == 32 ? 1 : 2);
    int idx = 0;
    size_t start = 0;
    for (;;) {
      const size_t semi = all.find(';', start);
      if (idx == want) {
        field = all.substr(start, semi == std::string::npos ? std::string::npos : semi - start);
        break;
      }
      if (semi == std::string::npos) break;
      start = semi + 1;
      ++idx;
    }
    if (!field.empty()) text = field;
  }
  MrfLaunchPlan pl;
  if (!ParseMrfPlan(text, &pl)) {
    std::fprintf(stderr, "[beatrice-b200] malformed MRF plan '%s' ignored\n", text.c_str());
    ParseMrfPlan("11|7|3", &pl);
  }
  return pl;
}

}  // namespace

void WaveState::Build(const WaveModel* m, int B_, int device_, TcMode tc) {
  AllocCond(m->dims, B_, device_);
  B = B_;
  device = device_;
  model = m;
  model_generation = m->generation;
  program.clear();
  arena.Clear();
  if (!ups_in_prologue) ups_in_prologue = std::make_shared<int>(0xE);
  fusable_ups_mask = 0;
  B200_CHECK(cudaSetDevice(device));
  const bool rc0 = m->dims.has_setter;
  const bool tcm = tc != kTcOff;
  const bool with_lo = tc == kTcSplit;

  // tensor-core mode: the pre conv runs as a one-block chain of the encoder cluster kernel (b200_enc.cu), which
  // keeps the six-row input history itself: cond then writes only the hop's fp32 row
  const bool pre_fused = tcm && m->pre_chain_ok && ResStackEnabled();
  static const bool no_cond_chain = [] {
    const char* ev = std::getenv("BEATRICE_B200_NO_COND_CHAIN");
    return ev && ev[0] == '1';
  }();
  const bool cond_fused = pre_fused && m->cond_chain_ok && !no_cond_chain;
  ring_hidden = arena.Plan(pre_fused ? 0 : spec::kPreK - 1, 1, kHidden);
  ring_pre = arena.Plan(1, 1, kHidden);
  // tensor-core mode: cond and pre also emit bf16 (hi [+ lo]) rings, so pre and ups0 move their input
  // with cp.async instead of gathering fp32 through registers
  const int ring_hidden_h = (tcm && !pre_fused) ? arena.PlanH(spec::kPreK - 1, 1, kHidden, with_lo) : -1;
  const int ring_pre_h = tcm ? arena.PlanH(1, 1, kHidden, with_lo) : -1;
  // fp32 rings: u (upsampler output), y (residual stream of each MRF branch), a (CUDA-core
  // path only).  Tensor-core mode keeps fp32 only where a residual / the branch sum needs it
  // (current rows) and adds bf16 rings -- already LeakyReLU'd -- that carry the conv history.
  int ring_u[4], ring_uh[4], ring_a[4][3][3], ring_y[4][3][4], ring_yh[4][3][4];
  // stages the fused MRF kernel covers (b200_mrf.cu): S streams per CTA, histories in mrf_hist
  bool fused[4] = {};
  int fused_S[4] = {}, fused_groups[4] = {}, fused_nc[4] = {};
  int t = 1;
  for (int s = 0; s < 4; ++s) {
    const int c = spec::kStageCh[s + 1];
    t *= spec::kRates[s];
    {
      // streams per CTA / cluster and cluster size (1 = the single-CTA kernel of b200_mrf.cu, > 1 = the
      // K-split cluster kernel of b200_mrfc.cu)
      int S = c >= 128 ? 16 : (c == 64 ? 6 : (c == 32 ? 3 : 1));
      int NC = c >= 128 ? 4 : 1;
      // developer overrides: BEATRICE_B200_MRF_S="s0,s1,s2,s3" (streams per CTA / cluster of stages 0..3),
      // BEATRICE_B200_MRF_NC="n0,n1,n2,n3" (cluster sizes), BEATRICE_B200_MRF_STAGES=bitmask of stages
      // allowed to use a fused kernel
      int mask = 0xF;
      if (const char* ev = std::getenv("BEATRICE_B200_MRF_STAGES")) mask = std::atoi(ev);
      if (const char* ev = std::getenv("BEATRICE_B200_MRF_S")) {
        int v[4] = {0, 0, 0, 0};
        if (std::sscanf(ev, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) == 4 && v[s] > 0) S = v[s];
      }
      if (const char* ev = std::getenv("BEATRICE_B200_MRF_NC")) {
        int v[4] = {0, 0, 0, 0};
        if (std::sscanf(ev, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) == 4 && v[s] > 0) NC = v[s];
      }
      // NC > 1 (stage 0): the clusters run k = 11 and k = 7, the k = 3 branch runs on the single-CTA kernel
      // beside them (16 CTAs on the SMs 32 clusters of 4 leave free), so the stage is one wave
      const bool form_ok = NC > 1 ? (MrfClusterSupported(c, NC, t, S, with_lo) && MrfFusedSupported(c, t, S, with_lo, spec::kMrfK[0]))
                                  : MrfFusedSupported(c, t, S, with_lo);
      fused[s] = tcm && FusedMrfEnabled() && ((mask >> s) & 1) && m->mrf_w_ptr[with_lo ? 1 : 0][s][0] != nullptr && form_ok;
      fused_S[s] = S;
      fused_nc[s] = NC;
      fused_groups[s] = (B + S - 1) / S;
    }
    if (fused[s]) {
      ring_u[s] = arena.Plan(0, t, c);
      ring_uh[s] = -1;
      for (int ki = 0; ki < 3; ++ki) {
        ring_y[s][ki][3] = arena.Plan(s < 3 ? 1 : spec::kPostK - 1, t, c);
        ring_stage_out[s][ki] = ring_y[s][ki][3];
      }
      continue;
    }
    const int u_hist = (spec::kMrfK[2] - 1) * spec::kMrfD[0];
    ring_u[s] = arena.Plan(tcm ? 0 : u_hist, t, c);
    ring_uh[s] = tcm ? arena.PlanH(u_hist, t, c, with_lo) : -1;
    for (int ki = 0; ki < 3; ++ki) {
      const int k = spec::kMrfK[ki];
      ring_y[s][ki][0] = ring_u[s];
      ring_yh[s][ki][0] = ring_uh[s];
      for (int di = 0; di < 3; ++di) {
        ring_a[s][ki][di] = tcm ? arena.PlanH(k - 1, t, c, with_lo) : arena.Plan(k - 1, t, c);
        const bool last = di == 2;
        const int hist = last ? (s < 3 ? 1 : spec::kPostK - 1) : (k - 1) * spec::kMrfD[di + 1];
        ring_y[s][ki][di + 1] = arena.Plan((tcm && !last) ? 0 : hist, t, c);
        ring_yh[s][ki][di + 1] = (tcm && !last) ? arena.PlanH(hist, t, c, with_lo) : -1;
      }
      ring_stage_out[s][ki] = ring_y[s][ki][3];
    }
  }
  arena.Commit(device, B);

  // fused-stage histories: one block per (stage, branch, conv), zero = silence
  size_t hist_off[4][3] = {};
  {
    size_t total = 0;
    std::vector<MrfHistBlock> blocks;
    for (int s = 0; s < 4; ++s) {
      if (!fused[s]) continue;
      for (int ki = 0; ki < 3; ++ki) {
        hist_off[s][ki] = total;
        total += (MrfHistElems(spec::kStageCh[s + 1], spec::kMrfK[ki], fused_S[s], fused_groups[s], with_lo) + 127) / 128 * 128;
      }
    }
    mrf_hist.Alloc(device, total * sizeof(uint16_t), true);
    static const int kDilPrefix[6] = {0, 1, 2, 5, 6, 11}, kDil[6] = {1, 1, 3, 1, 5, 1};
    for (int s = 0; s < 4; ++s) {
      if (!fused[s]) continue;
      const int c = spec::kStageCh[s + 1], P = with_lo ? 2 : 1;
      for (int ki = 0; ki < 3; ++ki) {
        const int k = spec::kMrfK[ki];
        const size_t unit = static_cast<size_t>(fused_groups[s]) * P * (c / 8) * fused_S[s] * 8 * (k - 1);
        for (int i = 0; i < 6; ++i) {
          MrfHistBlock hb;
          hb.base = mrf_hist.as<uint16_t>() + hist_off[s][ki] + unit * kDilPrefix[i];
          hb.planes_panels = P * (c / 8);
          hb.H = (k - 1) * kDil[i];
          hb.S = fused_S[s];
          hb.pad_ = 0;
          blocks.push_back(hb);
        }
      }
    }
    n_mrf_blocks = static_cast<int>(blocks.size());
    mrf_blocks.Alloc(device, sizeof(MrfHistBlock) * std::max<size_t>(blocks.size(), 1), true);
    if (!blocks.empty())
      UploadSync(mrf_blocks.p, blocks.data(), sizeof(MrfHistBlock) * blocks.size());
  }

  DescBuilder db;
  ConvDesc pre_d = MakeConv(arena.ring(ring_hidden), m->pre, 1, 1, 1, arena.ring(ring_pre), kActNone, kActNone);
  if (tcm && !pre_fused) {
    SetInH(&pre_d, arena.ring(ring_hidden_h));
    SetOutH(&pre_d, arena.ring(ring_pre_h), kActLrelu);   // ups0 reads lrelu(pre) as bf16
  }
  const int pre_idx = db.Add(pre_d);
  int ups_idx[4], c1_idx[4][3], c2_idx[4][3];
  t = 1;
  for (int s = 0; s < 4; ++s) {
    const int c = spec::kStageCh[s + 1];
    ConvDesc u = MakeConv(s == 0 ? arena.ring(ring_pre) : arena.ring(ring_stage_out[s - 1][0]), m->ups[s], 1, 1, t,
                          arena.ring(ring_u[s]), kActLrelu, kActNone);
    if (s > 0) {
      for (int ki = 0; ki < 3; ++ki) u.x[ki] = arena.ring(ring_stage_out[s - 1][ki]).base;
      u.n_x = 3;
      u.in_scale = 1.0f / 3.0f;
    }
    if (rc0 && !fused[s]) {   // fused stages apply the FiLM in the MRF kernel's prologue instead
      u.film = film[s].as<float>();
      u.film_C = c;
    }
    if (tcm && !fused[s]) SetOutH(&u, arena.ring(ring_uh[s]), kActLrelu);
    if (tcm && s == 0) SetInH(&u, arena.ring(ring_pre_h));
    ups_idx[s] = db.Add(u);
    t *= spec::kRates[s];
    for (int di = 0; di < 3 && !fused[s]; ++di) {
      for (int ki = 0; ki < 3; ++ki) {
        const Ring& a = arena.ring(ring_a[s][ki][di]);
        ConvDesc d1 = MakeConv(arena.ring(ring_y[s][ki][di]), m->c1[s][ki][di], spec::kMrfD[di], 1, t, a, kActLrelu,
                               kActLrelu);
        if (tcm) {
          SetInH(&d1, arena.ring(ring_yh[s][ki][di]));   // lrelu(y) with history, bf16
          SetOutH(&d1, a, kActNone);                      // v = lrelu(a) already
        }
        const int id = db.Add(d1);
        if (ki == 0) c1_idx[s][di] = id;
      }
      for (int ki = 0; ki < 3; ++ki) {
        const Ring& a = arena.ring(ring_a[s][ki][di]);
        ConvDesc d2 = MakeConv(a, m->c2[s][ki][di], 1, 1, t, arena.ring(ring_y[s][ki][di + 1]), kActNone, kActNone);
        SetRes(&d2, arena.ring(ring_y[s][ki][di]));
        if (tcm) {
          SetInH(&d2, a);
          if (ring_yh[s][ki][di + 1] >= 0) SetOutH(&d2, arena.ring(ring_yh[s][ki][di + 1]), kActLrelu);
        }
        const int id = db.Add(d2);
        if (ki == 0) c2_idx[s][di] = id;
      }
    }
  }
  const Ring out_ring = FlatRing(out.as<float>(), kOutHop, 1);
  ConvDesc pd = MakeConv(arena.ring(ring_stage_out[3][0]), m->post, 1, 1, kOutHop, out_ring, kActLrelu, kActTanh);
  for (int ki = 0; ki < 3; ++ki) pd.x[ki] = arena.ring(ring_stage_out[3][ki]).base;
  pd.n_x = 3;
  pd.in_scale = 1.0f / 3.0f;
  const int post_idx = db.Add(pd);

  descs.Alloc(device, sizeof(ConvDesc) * db.host.size(), false);
  UploadSync(descs.p, db.host.data(), sizeof(ConvDesc) * db.host.size());
  const ConvDesc* dd = descs.as<ConvDesc>();
  const int* frame = arena.frame();
  const int Bn = B;

  {
    const Ring hr = arena.ring(ring_hidden);
    const Ring hh = (tcm && !pre_fused) ? arena.ring(ring_hidden_h) : Ring();
    const float* ph = phone_in.as<float>();
    const int* q = q_in.as<int>();
    const float* ft = feat_in.as<float>();
    const float* sp = spk.as<float>();
    const float* fm = rc0 ? formant.as<float>() : nullptr;
    const WaveModel* mm = m;
    Op op;
    op.name = "wave.cond";
    op.flops = 2.0 * B * (m->dims.phone_channels + kPitchFeatures) * kHidden;
    op.bytes = 4.0 * (m->dims.phone_channels * kHidden + B * (m->dims.phone_channels + 4 * kHidden));
    op.launch = [=](cudaStream_t s) {
      LaunchCond(ph, mm->dims.phone_channels, q, mm->dims.pitch_bins, ft, mm->embed.w, mm->embed.b, mm->pitch_emb,
                 mm->feat_proj, sp, fm, hr.base, hh.hi, hh.lo, hr.slots, Bn, frame, s);   // hh has hr's geometry
    };
    if (!cond_fused) program.push_back(op);   // fused: first block of the "wave.cond+pre" chain below
  }
  auto add_gemm = [&](const std::string& name, int idx, int nz, bool mrf) {
    Op op;
    op.name = name;
    for (int z = 0; z < nz; ++z) {
      op.flops += ConvFlops(db.host[idx + z], B);
      op.bytes += ConvBytes(db.host[idx + z], B);
    }
    op.is_mrf = mrf;
    const ConvDesc h = Traced(db.host[idx], name);
    const ConvDesc* dp = dd + idx;
    op.launch = GemmLauncher(dp, h, nz, Bn, frame, tc, true);
    program.push_back(op);
  };
  n_pre_blocks = 0;
  if (pre_fused) {
    const int d3 = 3;
    pre_hist.Alloc(device, ResStackHistElems(kHidden, 1, &d3, B) * sizeof(uint16_t), true);
    std::vector<MrfHistBlock> blocks;
    ResStackHistBlocks(kHidden, 1, &d3, B, pre_hist.as<uint16_t>(), &blocks);
    n_pre_blocks = static_cast<int>(blocks.size());
    pre_blocks.Alloc(device, sizeof(MrfHistBlock) * blocks.size(), false);
    UploadSync(pre_blocks.p, blocks.data(), sizeof(MrfHistBlock) * blocks.size());
    ResStackParams rp;
    std::memset(&rp, 0, sizeof(rp));
    rp.x_in = arena.ring(ring_hidden).base;             // [B][256]: the hop's conditioning row (one slot)
    const Ring& pr = arena.ring(ring_pre);
    const Ring& ph = arena.ring(ring_pre_h);
    rp.x_out = pr.base;
    rp.xh_out = ph.hi;
    rp.xl_out = ph.lo;
    rp.out_slots = pr.slots;                            // == ph.slots: ups0 reads one row of history
    rp.out_act = 1;                                     // ups0 consumes lrelu(pre) as bf16
    rp.frame = frame;
    rp.w = m->pre_w.as<uint16_t>();
    rp.bias = m->pre_par.as<float>();
    rp.gamma = rp.bias + kHidden;
    rp.beta = rp.gamma + kHidden;
    rp.hist = pre_hist.as<uint16_t>();
    rp.n_blk = 1;
    rp.kind[0] = 3;
    rp.dil[0] = 1;
    rp.B = B;
    rp.n_tiles = ResStackTiles(B);
    Op op;
    op.name = "wave.pre";
    op.flops = ConvFlops(db.host[pre_idx], B);
    op.bytes = ConvBytes(db.host[pre_idx], B);
    if (cond_fused) {   // conditioning + pre conv as one chain: no cond launch, no hidden ring
      rp.x_in = phone_in.as<float>();
      rp.x_in_C = m->dims.phone_channels;
      rp.emb_q = q_in.as<int>();
      rp.emb_bins = m->dims.pitch_bins;
      rp.emb_pitch = m->pitch_emb;
      rp.emb_feat = feat_in.as<float>();
      rp.emb_wf = m->feat_proj;
      rp.emb_spk = spk.as<float>();
      rp.emb_formant = rc0 ? formant.as<float>() : nullptr;
      rp.w = m->cond_pre_w.as<uint16_t>();
      rp.bias = m->cond_pre_par.as<float>();
      rp.gamma = rp.bias + 2 * kHidden;
      rp.beta = rp.gamma + 2 * kHidden;
      rp.n_blk = 2;
      rp.kind[0] = 4;
      rp.kind[1] = 3;
      rp.dil[1] = 1;
      op.name = "wave.cond+pre";
      op.flops += 2.0 * B * (m->dims.phone_channels + kPitchFeatures) * kHidden;
    }
    op.launch = [=](cudaStream_t s) { LaunchResStack(rp, kHidden, s); };
    program.push_back(op);
  } else {
    add_gemm("wave.pre", pre_idx, 1, false);
  }
  int t_stage = 1;
  // developer switch: BEATRICE_B200_FUSE_UPS=0 keeps every upsampler a launch of its own
  static const bool fuse_ups_enabled = [] {
    const char* ev = std::getenv("BEATRICE_B200_FUSE_UPS");
    return !(ev && ev[0] == '0');
  }();
  for (int s = 0; s < 4; ++s) {
    // stages 1..3 on the single-CTA fused kernel compute their upsampler in the kernel's prologue (MrfUpsDesc)
    const bool ups_fused = fuse_ups_enabled && s >= 1 && fused[s] && fused_nc[s] == 1 && with_lo && m->ups_img_ptr[s] != nullptr &&
                           MrfUpsFusable(spec::kStageCh[s + 1], t_stage * spec::kRates[s], fused_S[s], spec::kRates[s], with_lo);
    add_gemm("wave.ups" + std::to_string(s), ups_idx[s], 1, false);
    if (ups_fused) {   // launches only while the prologue form is switched off (see WaveState::ups_in_prologue)
      const std::function<void(cudaStream_t)> plain = program.back().launch;
      const std::shared_ptr<int> in_prologue = ups_in_prologue;
      program.back().launch = [=](cudaStream_t st) {
        if (!((*in_prologue >> s) & 1)) plain(st);
      };
      fusable_ups_mask |= 1 << s;
    }
    t_stage *= spec::kRates[s];
    if (fused[s]) {
      const int c = spec::kStageCh[s + 1];
      MrfStageParams mp;
      std::memset(&mp, 0, sizeof(mp));
      for (int ki = 0; ki < 3; ++ki) {
        const Ring& o = arena.ring(ring_stage_out[s][ki]);
        mp.br[ki].w = m->mrf_w_ptr[with_lo ? 1 : 0][s][ki];
        mp.br[ki].bias = m->mrf_bias_ptr[s][ki];
        mp.br[ki].hist = mrf_hist.as<uint16_t>() + hist_off[s][ki];
        mp.br[ki].out = o.base;
        mp.br[ki].out_slots = o.slots;
        mp.br[ki].k = spec::kMrfK[ki];
      }
      const Ring& ur = arena.ring(ring_u[s]);
      mp.u = ur.base;
      mp.film = rc0 ? film[s].as<float>() : nullptr;
      mp.u_slots = ur.slots;
      mp.T = t_stage;
      mp.S = fused_S[s];
      mp.MT = (fused_S[s] * t_stage + 127) / 128;
      mp.B = B;
      mp.n_groups = fused_groups[s];
      mp.frame = frame;
      MrfUpsDesc ups_desc;
      std::memset(&ups_desc, 0, sizeof(ups_desc));
      if (ups_fused) {
        ups_desc.w = m->ups_img_ptr[s];
        ups_desc.bias = m->ups_bias_raw[s];
        for (int ki = 0; ki < 3; ++ki) ups_desc.x[ki] = arena.ring(ring_stage_out[s - 1][ki]).base;
        ups_desc.x_slots = arena.ring(ring_stage_out[s - 1][0]).slots;
        ups_desc.r = spec::kRates[s];
      }
      mp.n_branches = 3;
      // CTA scheduling order of the branches.  Where the stage fits in one wave: longest chain (k = 11) first.
      // Stage 3 (768 short CTAs for 592 slots): k = 11, then k = 3, then k = 7 -- the CTAs that have to wait
      // for a slot then start when the k = 3 CTAs retire (~1/4 of the stage) instead of when the k = 7 ones do.
      mp.y2br[0] = 2;
      mp.y2br[1] = c <= 16 ? 0 : 1;
      mp.y2br[2] = c <= 16 ? 1 : 0;
      mp.pdl_mode = 0;
      if (const char* ev = std::getenv("BEATRICE_B200_MRF_TRACE")) mp.trace = std::atoi(ev);
      // developer switch: BEATRICE_B200_MRF_LATE=<stage bit mask>
      if (const char* ev = std::getenv("BEATRICE_B200_MRF_LATE")) mp.late_launch = (std::atoi(ev) >> s) & 1;
      Op op;
      op.name = "wave.mrf" + std::to_string(s) + ".fused";
      op.flops = 2.0 * c * c * t_stage * B * 6.0 * (3 + 7 + 11);
      op.bytes = 2.0 * 6 * 21 * c * c + 4.0 * B * t_stage * c * 4;

      op.is_mrf = true;
      const int nc = fused_nc[s];
      if (nc > 1) {
        MrfStageParams mc = mp, m3 = mp;
        mc.n_branches = 2;      // blockIdx.y = 0 -> k = 11, 1 -> k = 7
        m3.n_branches = 1;      // k = 3 on the single-CTA kernel ...
        m3.y2br[0] = 0;
        MrfOneBranchPerCta(&m3);
        m3.pdl_mode = 1;        // ... as the second launch of the pair (see MrfStageParams::pdl_mode)
        op.launch = [=](cudaStream_t st) {
          LaunchMrfStageCluster(mc, c, nc, with_lo, st);
          LaunchMrfStage(m3, c, with_lo, st);
        };
      } else {
        // launches / CTA classes of the stage (MrfLaunchPlan): every launch after the first is the "second launch of
        // a pair" with respect to the one before it, so the last one completing implies the stage is complete
        MrfLaunchPlan plan = MrfPlanFor(c);
        for (int attempt = 0; attempt < 2; ++attempt) {
          bool ok = true;
          for (int li = 0; li < plan.n_launch; ++li) {
            int kmax = 3, nbm = 1;
            for (int y = 0; y < plan.l[li].n_y; ++y) {
              nbm = std::max(nbm, plan.l[li].ylen[y]);
              for (int bi = 0; bi < plan.l[li].ylen[y]; ++bi) kmax = std::max(kmax, spec::kMrfK[plan.l[li].yseq[y][bi]]);
            }
            ok = ok && MrfFusedSupported(c, t_stage, fused_S[s], with_lo, kmax, nbm, ups_fused);
          }
          if (ok) break;
          ParseMrfPlan("11|7|3", &plan);
        }
        std::vector<MrfStageParams> launches;
        for (int li = 0; li < plan.n_launch; ++li) {
          MrfStageParams ml = mp;
          const MrfLaunchPlan::L& L = plan.l[li];
          ml.n_branches = L.n_y;
          ml.nb_max = 1;
          for (int y = 0; y < 3; ++y) {
            ml.ylen[y] = y < L.n_y ? L.ylen[y] : 0;
            ml.nb_max = std::max(ml.nb_max, ml.ylen[y]);
            for (int bi = 0; bi < 3; ++bi) ml.yseq[y][bi] = L.yseq[y][bi];
            ml.y2br[y] = L.yseq[y][0];
          }
          ml.smem_min = L.smem_kb * 1024;
          ml.pdl_mode = li == 0 ? 0 : 1;
          launches.push_back(ml);
        }
        std::vector<MrfStageParams> launches_ups = launches;   // the same launches with the upsampler in their prologue
        for (MrfStageParams& ml : launches_ups) ml.ups = ups_desc;
        const std::shared_ptr<int> in_prologue = ups_in_prologue;
        op.launch = [=](cudaStream_t st) {
          for (const MrfStageParams& ml : (ups_fused && ((*in_prologue >> s) & 1)) ? launches_ups : launches) LaunchMrfStage(ml, c, with_lo, st);
        };
      }
      program.push_back(op);
      continue;
    }
    for (int di = 0; di < 3; ++di) {
      add_gemm("wave.mrf" + std::to_string(s) + ".d" + std::to_string(spec::kMrfD[di]) + ".c1", c1_idx[s][di], 3, true);
      add_gemm("wave.mrf" + std::to_string(s) + ".d" + std::to_string(spec::kMrfD[di]) + ".c2", c2_idx[s][di], 3, true);
    }
  }
  {
    const ConvDesc h = db.host[post_idx];
    const ConvDesc* dp = dd + post_idx;
    Op op;
    op.name = "wave.post";
    op.flops = ConvFlops(h, B);
    op.bytes = ConvBytes(h, B);
    AdvanceFold fold;
    advance_folded = fold_frames[0] != nullptr && fold_frames[1] != nullptr && PostConvFused(h);
    if (advance_folded) {
      fold_done.Alloc(device, sizeof(int), true);
      fold.done = fold_done.as<int>();
      fold.frames[0] = arena.frame();
      fold.frames[1] = fold_frames[0];
      fold.frames[2] = fold_frames[1];
    }
    op.launch = [=](cudaStream_t s) { LaunchPostConv(dp, h, Bn, frame, fold, s); };
    program.push_back(op);
    // the same launch advancing this state's counter only: a pipeline drain vocodes a hop the encoders took earlier
    AdvanceFold own = fold;
    own.frames[1] = own.frames[2] = nullptr;
    post_own_advance = [=](cudaStream_t s) { LaunchPostConv(dp, h, Bn, frame, own, s); };
  }
  if (!advance_folded) {
    int* f = arena.frame();
    Op op;
    op.name = "wave.advance";
    op.launch = [=](cudaStream_t s) { LaunchAdvance(f, s); };
    program.push_back(op);
  }
}

void WaveState::ZeroAll(cudaStream_t s) {
  arena.ZeroAll(s);
  if (mrf_hist.p) B200_CHECK(cudaMemsetAsync(mrf_hist.p, 0, mrf_hist.bytes, s));
  if (n_pre_blocks > 0 && pre_hist.p) B200_CHECK(cudaMemsetAsync(pre_hist.p, 0, pre_hist.bytes, s));
}
void WaveState::ZeroStream(int b, cudaStream_t s) {
  arena.ZeroStream(b, s);
  LaunchMrfZeroStream(mrf_blocks.as<MrfHistBlock>(), n_mrf_blocks, b, s);
  LaunchMrfZeroStream(pre_blocks.as<MrfHistBlock>(), n_pre_blocks, b, s);
}

void RunProgram(const std::vector<Op>& program, cudaStream_t s) {
  for (const Op& op : program) op.launch(s);
}

GraphRunner::~GraphRunner() { Reset(); }
void GraphRunner::Reset() {
  if (exec_) cudaGraphExecDestroy(exec_);
  exec_ = nullptr;
}
void GraphRunner::Run(cudaStream_t s, const std::function<void(cudaStream_t)>& body, bool use_graph) {
  if (!use_graph) {
    body(s);
    return;
  }
  if (!exec_) {
    cudaGraph_t graph = nullptr;
    B200_CHECK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    try {
      body(s);
    } catch (...) {   // leave capture mode before the failure travels up to the ABI boundary
      cudaStreamEndCapture(s, &graph);
      if (graph) cudaGraphDestroy(graph);
      (void)cudaGetLastError();
      throw;
    }
    B200_CHECK(cudaStreamEndCapture(s, &graph));
    B200_CHECK(cudaGraphInstantiate(&exec_, graph, 0));
    B200_CHECK(cudaGraphDestroy(graph));
  }
  B200_CHECK(cudaGraphLaunch(exec_, s));
}

}  // namespace b200
