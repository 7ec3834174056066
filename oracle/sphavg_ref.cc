// TEST INFRASTRUCTURE ONLY -- known answers for the voice-morphing averages, straight from the reference's
// src/common/spherical_average.h (included where it lies under /root/reference; nothing is copied).
//
//   sphavg_ref <M: 128|256> <n_speakers> <rows> <points.f32> <weights.f32> <out.f32>
//
// points.f32: [n_speakers][rows][M] (the model's additive table with rows = 1, its key-value table with rows = 384);
// weights.f32: 256 floats, the argument of ProcessorCore2::SetSpeakerMorphingWeights.  The weights are prepared,
// arg-sorted and pruned to the 8 heaviest exactly as ApplySpeakerMorphingWeights does (processor_core_2.cc:507-532,
// voice_morph_state.h:87-104 -- both compiled from the reference), then for every row: Initialize(n, M, points of
// that row, min(n, 8)), SetWeights, at most kSphAvgMaxNUpdates = 4 Update(), GetResult -- the per-frame sequence of
// processor_core_2.cc:123-165.  out.f32: [rows][M].
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "common/model_config.h"
#include "common/spherical_average.h"
#include "common/voice_morph_state.h"

using beatrice::common::kMaxNSpeakers;

template <std::size_t M>
int Run(int n, int rows, const std::vector<float>& pts, const std::array<float, kMaxNSpeakers>& raw, const char* out_path) {
  const auto weights = beatrice::common::PrepareVoiceMorphWeights(raw, n);
  std::array<int, kMaxNSpeakers> indices{};
  std::iota(indices.data(), indices.data() + n, 0);
  std::sort(indices.data(), indices.data() + n, [&weights](const int a, const int b) -> bool { return weights[a] > weights[b]; });
  std::array<float, kMaxNSpeakers> pruned{};
  const int n_weights = std::min(n, 8);
  for (int i = 0; i < n_weights; ++i) pruned[indices[i]] = weights[indices[i]];
  std::vector<float> out(static_cast<size_t>(rows) * M);
  beatrice::common::AlignedVector<float, 64> block(static_cast<size_t>(n) * M), dst(M);
  for (int r = 0; r < rows; ++r) {
    for (int j = 0; j < n; ++j) std::memcpy(&block[j * M], &pts[(static_cast<size_t>(j) * rows + r) * M], sizeof(float) * M);
    beatrice::common::SphericalAverage<float, M> avg;
    avg.Initialize(n, M, block.data(), std::min(n, 8));
    avg.SetWeights(n, pruned.data(), indices.data());
    for (int j = 0; j < 4; ++j)
      if (avg.Update()) break;
    avg.GetResult(M, dst.data());
    std::memcpy(&out[static_cast<size_t>(r) * M], dst.data(), sizeof(float) * M);
  }
  FILE* f = std::fopen(out_path, "wb");
  if (!f) return 5;
  std::fwrite(out.data(), 4, out.size(), f);
  std::fclose(f);
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 7) {
    std::fprintf(stderr, "usage: sphavg_ref <M> <n_speakers> <rows> <points.f32> <weights.f32> <out.f32>\n");
    return 2;
  }
  const int M = std::atoi(argv[1]), n = std::atoi(argv[2]), rows = std::atoi(argv[3]);
  std::vector<float> pts(static_cast<size_t>(n) * rows * M);
  std::array<float, kMaxNSpeakers> raw{};
  FILE* f = std::fopen(argv[4], "rb");
  if (!f || std::fread(pts.data(), 4, pts.size(), f) != pts.size()) return 3;
  std::fclose(f);
  f = std::fopen(argv[5], "rb");
  if (!f || std::fread(raw.data(), 4, raw.size(), f) != raw.size()) return 4;
  std::fclose(f);
  if (M == 128) return Run<128>(n, rows, pts, raw, argv[6]);
  if (M == 256) return Run<256>(n, rows, pts, raw, argv[6]);
  return 2;
}
