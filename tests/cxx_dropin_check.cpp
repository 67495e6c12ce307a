// Compile/link/run check of include/richdem_b200.hpp against the reference headers.
// Existing RichDEM call sites (float / double rasters) resolve to the B200 explicit specialisations
// with no source change; the same templates instantiated for `double` elevations still run the
// reference's CPU code, which this program uses as the in-process oracle (float values are exactly
// representable in double, and fill / D8 directions / D8 accumulation only compare elevations).
// Without a GPU every B200 call must throw std::runtime_error (there is no CPU fallback).
#include <richdem_b200.hpp>

#include <cstdio>
#include <cstdlib>

using namespace richdem;

static unsigned lcg(unsigned &s) { return s = s * 1664525u + 1013904223u; }

int main() {
  const int W = 300, H = 217;
  Array2D<float> dem(W, H, 0.f);
  dem.setNoData(-9999.f);
  unsigned seed = 12345;
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      const float ridge = 40.f * ((x / 37 + y / 29) % 3);
      dem(x, y) = ridge + (float)(lcg(seed) >> 24) * 0.5f + 0.05f * (float)((x * 3 + y * 5) % 17);
    }
  Array2D<double> demd(dem);  // templateCopy + element-wise copy
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) demd(x, y) = dem(x, y);
  demd.setNoData(-9999.0);

  int thrown = 0, calls = 0, mism = 0;
  auto attempt = [&](const char *name, auto &&fn) {
    calls++;
    try {
      fn();
      std::printf("%-24s ran\n", name);
      return true;
    } catch (const std::runtime_error &e) {
      thrown++;
      std::printf("%-24s runtime_error: %s\n", name, e.what());
      return false;
    }
  };
  auto compare = [&](const char *name, auto &a, auto &b) {
    long bad = 0;
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++)
        if ((double)a(x, y) != (double)b(x, y)) bad++;
    std::printf("  %-22s %s (%ld cells differ)\n", name, bad ? "MISMATCH" : "identical", bad);
    if (bad) mism++;
  };

  // fill: B200 (float specialisation) vs reference CPU (double instantiation)
  if (attempt("FillDepressions<D8>", [&] { FillDepressions<Topology::D8>(dem); })) {
    FillDepressions<Topology::D8>(demd);
    compare("fill", dem, demd);
  }
  Array2D<uint8_t> dirs, dirsd;
  if (attempt("d8_flow_directions", [&] { d8_flow_directions(dem, dirs); })) {
    d8_flow_directions(demd, dirsd);
    compare("d8 directions", dirs, dirsd);
  }
  Array2D<int32_t> area;
  if (attempt("d8_flow_accum", [&] { if (dirs.width() != W) dirs.resize(dem); d8_flow_accum(dirs, area); })) {
    Array2D<int32_t> aread;
    Array2D<int32_t> dirs32(dirsd);
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) dirs32(x, y) = dirsd(x, y);
    dirs32.setNoData(255);
    d8_flow_accum(dirs32, aread);  // <int32,int32>: the reference's CPU code
    compare("d8 flow accum", area, aread);
  }
  Array2D<double> accum(dem, 1.0), accumd(dem, 1.0);
  if (attempt("FA_D8", [&] { FA_D8(dem, accum); })) {
    FA_D8(demd, accumd);
    compare("FA_D8", accum, accumd);
  }
  Array3D<float> props(dem), propsd(demd);
  if (attempt("FM_D8", [&] { FM_D8(dem, props); })) {
    FM_D8(demd, propsd);
    long bad = 0;
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++)
        for (int n = 0; n < 9; n++)
          if (props(x, y, n) != propsd(x, y, n)) bad++;
    std::printf("  %-22s %s (%ld slots differ)\n", "FM_D8", bad ? "MISMATCH" : "identical", bad);
    if (bad) mism++;
  }
  Array2D<double> acc2(dem, 1.0);
  attempt("FlowAccumulation(props)", [&] { FlowAccumulation(props, acc2); });
  if (thrown == 0) compare("FlowAccumulation", acc2, accumd);
  attempt("ResolveFlatsEpsilon", [&] { ResolveFlatsEpsilon(dem); });
  Array2D<double> acc3(dem, 1.0);
  attempt("FA_Tarboton", [&] { FA_Tarboton(dem, acc3); });
  void (*fp)(const Array2D<float> &, Array2D<double> &) = &FA_D8<float, double>;  // pyrichdem-style binding
  Array2D<double> acc4(dem, 1.0);
  attempt("&FA_D8<float,double>", [&] { fp(dem, acc4); });
  // error convention: dimension mismatch is a std::runtime_error (generic.hpp:42-43)
  bool dim_throw = false;
  try {
    Array2D<double> small(3, 3, 1.0);
    FA_D8(dem, small);
  } catch (const std::runtime_error &) {
    dim_throw = true;
  }
  std::printf("calls=%d thrown=%d mismatches=%d dim_mismatch_throws=%d\n", calls, thrown, mism, (int)dim_throw);
  if (thrown != 0 && thrown != calls) return 2;  // partial failure
  return (mism == 0 && dim_throw) ? 0 : 1;
}
