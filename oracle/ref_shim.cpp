// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from the product path.
//
// Thin extern "C" shim that instantiates the UNMODIFIED reference templates from
// /root/reference/include (header-only; nothing is copied into this repo) so tests can
// call the real RichDEM CPU implementation through ctypes.  Built by oracle/Makefile into
// oracle/_ref/libref_richdem.so (git-ignored, travels to the GPU box with the snapshot).
//
// Every entry point takes plain row-major host buffers (i = y*W + x), wraps them unowned in
// richdem::Array2D (reference: include/richdem/common/Array2D.hpp:344-352) and calls the
// reference function named in the comment beside it.
#include <richdem/common/Array2D.hpp>
#include <richdem/common/Array3D.hpp>
#include <richdem/depressions/depressions.hpp>
#include <richdem/depressions/Zhou2016.hpp>
#include <richdem/depressions/Barnes2014.hpp>
#include <richdem/flats/flats.hpp>
#include <richdem/flats/find_flats.hpp>
#include <richdem/flats/flat_resolution.hpp>
#include <richdem/flowmet/d8_flowdirs.hpp>
#include <richdem/methods/d8_methods.hpp>
#include <richdem/methods/flow_accumulation.hpp>
#include <richdem/methods/terrain_attributes.hpp>

#include <omp.h>

#include <cstdint>
#include <cstring>

using namespace richdem;

extern "C" {

// depressions/depressions.hpp:13-21 -> depressions/Zhou2016.hpp:125-191
void ref_fill_depressions_d8_f32(float *dem, int w, int h) {
  Array2D<float> a(dem, w, h);
  FillDepressions<Topology::D8>(a);
}

// depressions/depressions.hpp:16-17 -> depressions/Barnes2014.hpp:230-304 with the 4-neighbour topology
void ref_fill_depressions_d4_f32(float *dem, int w, int h) {
  Array2D<float> a(dem, w, h);
  FillDepressions<Topology::D4>(a);
}

// depressions/Barnes2014.hpp:336-420 (pyrichdem FillDepressions(epsilon=True)); not on the B200 path -- kept in the checker to
// show what its result depends on (tests/test_oracle.py::test_epsilon_fill_depends_on_the_queue_order)
void ref_priority_flood_epsilon_f32(float *dem, int w, int h, float nodata) {
  Array2D<float> a(dem, w, h);
  a.setNoData(nodata);
  PriorityFloodEpsilon_Barnes2014<Topology::D8>(a);
}

// depressions/Zhou2016.hpp:125-191 (the function pyrichdem binds, pywrapper.hpp:32)
void ref_priority_flood_zhou2016_f32(float *dem, int w, int h) {
  Array2D<float> a(dem, w, h);
  PriorityFlood_Zhou2016(a);
}

// depressions/Barnes2014.hpp:230-304 (independent second oracle)
void ref_priority_flood_barnes2014_f32(float *dem, int w, int h) {
  Array2D<float> a(dem, w, h);
  PriorityFlood_Barnes2014<Topology::D8>(a);
}

// depressions/Barnes2014.hpp:136-198
void ref_priority_flood_original_f32(float *dem, int w, int h) {
  Array2D<float> a(dem, w, h);
  PriorityFlood_Original<Topology::D8>(a);
}

// flats/find_flats.hpp:28-69
void ref_find_flats_f32(const float *dem, int w, int h, float nodata, int8_t *flats) {
  Array2D<float> a(const_cast<float *>(dem), w, h);
  a.setNoData(nodata);
  Array2D<int8_t> f;
  FindFlats(a, f);
  std::memcpy(flats, f.data(), (size_t)w * h);
}

// flats/Barnes2014.hpp:398-467 ; labels are order-dependent ids (only the partition matters)
void ref_get_flat_mask_f32(const float *dem, int w, int h, float nodata, int32_t *mask,
                           int32_t *labels) {
  Array2D<float> a(const_cast<float *>(dem), w, h);
  a.setNoData(nodata);
  Array2D<int32_t> m, l;
  GetFlatMask(a, m, l);
  std::memcpy(mask, m.data(), sizeof(int32_t) * (size_t)w * h);
  std::memcpy(labels, l.data(), sizeof(int32_t) * (size_t)w * h);
}

// flats/flats.hpp:21-28
void ref_resolve_flats_epsilon_f32(float *dem, int w, int h, float nodata) {
  Array2D<float> a(dem, w, h);
  a.setNoData(nodata);
  ResolveFlatsEpsilon(a);
}

// flowmet/d8_flowdirs.hpp:96-123
void ref_d8_flow_directions_f32(const float *dem, int w, int h, float nodata, uint8_t *dirs) {
  Array2D<float> a(const_cast<float *>(dem), w, h);
  a.setNoData(nodata);
  Array2D<uint8_t> d;
  d8_flow_directions(a, d);
  std::memcpy(dirs, d.data(), (size_t)w * h);
}

// methods/d8_methods.hpp:47-139  (direction grid NoData = 255, constants.hpp:76)
void ref_d8_flow_accum_u8_i32(const uint8_t *dirs, int w, int h, int32_t *area) {
  Array2D<uint8_t> d(const_cast<uint8_t *>(dirs), w, h);
  d.setNoData(FLOWDIR_NO_DATA);
  Array2D<int32_t> a;
  // the reference's dependency pass increments shared counters from an OpenMP loop without
  // synchronisation (d8_methods.hpp:68-92); run it single-threaded so the oracle is deterministic
  const int nt = omp_get_max_threads();
  omp_set_num_threads(1);
  d8_flow_accum(d, a);
  omp_set_num_threads(nt);
  std::memcpy(area, a.data(), sizeof(int32_t) * (size_t)w * h);
}

// same function on an int32 direction grid with an explicit NoData (the tests/flow_accum
// fixtures are loaded as int32 with NODATA_value -1, tests/tests.cpp:135-146)
void ref_d8_flow_accum_i32_i32(const int32_t *dirs, int w, int h, int32_t nodata, int32_t *area) {
  Array2D<int32_t> d(const_cast<int32_t *>(dirs), w, h);
  d.setNoData(nodata);
  Array2D<int32_t> a;
  const int nt = omp_get_max_threads();
  omp_set_num_threads(1);
  d8_flow_accum(d, a);
  omp_set_num_threads(nt);
  std::memcpy(area, a.data(), sizeof(int32_t) * (size_t)w * h);
}

// flowmet/OCallaghan1984.hpp:81-84 ; props is [y][x][9] float (Array3D.hpp:203-206)
void ref_fm_d8_f32(const float *dem, int w, int h, float nodata, float *props) {
  Array2D<float> a(const_cast<float *>(dem), w, h);
  a.setNoData(nodata);
  Array3D<float> p(props, w, h);  // unowned wrap, Array3D.hpp:118
  FM_D8(a, p);
}

// flowmet/Tarboton1997.hpp:14-144
void ref_fm_tarboton_f32(const float *dem, int w, int h, float nodata, float *props) {
  Array2D<float> a(const_cast<float *>(dem), w, h);
  a.setNoData(nodata);
  Array3D<float> p(props, w, h);
  FM_Tarboton(a, p);
}

// methods/flow_accumulation_generic.hpp:33-100 ; props NoData = -2 (constants.hpp:85)
void ref_flow_accumulation_props_f64(const float *props, int w, int h, double *accum) {
  Array3D<float> p(const_cast<float *>(props), w, h);
  p.setNoData(NO_DATA_GEN);
  Array2D<double> a(accum, w, h);
  FlowAccumulation(p, a);
}

// methods/flow_accumulation.hpp:27
void ref_fa_d8_f32_f64(const float *dem, int w, int h, float nodata, double *accum) {
  Array2D<float> a(const_cast<float *>(dem), w, h);
  a.setNoData(nodata);
  Array2D<double> acc(accum, w, h);
  FA_D8(a, acc);
}

// methods/flow_accumulation.hpp:16
void ref_fa_tarboton_f32_f64(const float *dem, int w, int h, float nodata, double *accum) {
  Array2D<float> a(const_cast<float *>(dem), w, h);
  a.setNoData(nodata);
  Array2D<double> acc(accum, w, h);
  FA_Tarboton(a, acc);
}

// method numbering of the C ABI: 0 FM_D8, 1 FM_Tarboton, 2 FM_D4 (flowmet/OCallaghan1984.hpp:89-91),
// 3 FM_Holmgren (flowmet/Holmgren1994.hpp:13-83; xparam 1 is what FM_Quinn passes, Quinn1991.hpp:15),
// 4 FM_Freeman (flowmet/Freeman1991.hpp:13-80)
void ref_fm_method_f32(int method, const float *dem, int w, int h, float nodata, double xparam, float *props) {
  Array2D<float> a(const_cast<float *>(dem), w, h);
  a.setNoData(nodata);
  Array3D<float> p(props, w, h);
  switch (method) {
    case 0: FM_D8(a, p); break;
    case 1: FM_Tarboton(a, p); break;
    case 2: FM_D4(a, p); break;
    case 3: if (xparam == 1.0) FM_Quinn(a, p); else FM_Holmgren(a, p, xparam); break;
    default: FM_Freeman(a, p, xparam); break;
  }
}

// methods/flow_accumulation.hpp:28,19,18,20
void ref_fa_method_f32_f64(int method, const float *dem, int w, int h, float nodata, double xparam, double *accum) {
  Array2D<float> a(const_cast<float *>(dem), w, h);
  a.setNoData(nodata);
  Array2D<double> acc(accum, w, h);
  switch (method) {
    case 0: FA_D8(a, acc); break;
    case 1: FA_Tarboton(a, acc); break;
    case 2: FA_D4(a, acc); break;
    case 3: if (xparam == 1.0) FA_Quinn(a, acc); else FA_Holmgren(a, acc, xparam); break;
    default: FA_Freeman(a, acc, xparam); break;
  }
}

// methods/terrain_attributes.hpp:370-538; attribute numbering of the C ABI (0 slope rise/run, 1 percentage, 2 degrees,
// 3 radians, 4 aspect, 5 curvature, 6 planform, 7 profile).  The output raster keeps its own NoData (TerrainProcessor
// :344 resizes without copying it), which is what nodata_out stands for.
void ref_terrain_attribute_f32(int attribute, const float *dem, int w, int h, float nodata_in, float nodata_out, float zscale,
                               double cell_x, double cell_y, float *out) {
  Array2D<float> a(const_cast<float *>(dem), w, h);
  a.setNoData(nodata_in);
  a.geotransform = {0.0, cell_x, 0.0, 0.0, 0.0, -cell_y};
  Array2D<float> o;
  o.setNoData(nodata_out);
  switch (attribute) {
    case 0: TA_slope_riserun(a, o, zscale); break;
    case 1: TA_slope_percentage(a, o, zscale); break;
    case 2: TA_slope_degrees(a, o, zscale); break;
    case 3: TA_slope_radians(a, o, zscale); break;
    case 4: TA_aspect(a, o, zscale); break;
    case 5: TA_curvature(a, o, zscale); break;
    case 6: TA_planform_curvature(a, o, zscale); break;
    default: TA_profile_curvature(a, o, zscale); break;
  }
  std::memcpy(out, o.data(), sizeof(float) * (size_t)w * h);
}

// common/Array2D.hpp:209-241 (saveToCache) / :246-281 (loadNative through the `native` constructor, :420-423): the
// reference's own cache format, written and read by the reference, so that the Python layer's SaveNative / LoadNative can
// be checked against it in both directions.  geotransform6: the six doubles; projection: a C string.
void ref_save_native_f32(const char *path, const float *data, int w, int h, float nodata, const double *geotransform6,
                         const char *projection) {
  Array2D<float> a(w, h, 0.0f);
  std::memcpy(a.data(), data, sizeof(float) * (size_t)w * h);
  a.setNoData(nodata);
  a.geotransform.assign(geotransform6, geotransform6 + 6);
  a.projection = projection ? projection : "";
  a.saveToCache(path);
}
// returns 0 and fills the outputs (data may be null to read the header only); dims[0..1] = width, height
int ref_load_native_f32(const char *path, float *data, int *dims, float *nodata, double *geotransform6, char *projection,
                        int projection_capacity) {
  Array2D<float> a(std::string(path), true);
  dims[0] = a.width();
  dims[1] = a.height();
  *nodata = a.noData();
  for (int k = 0; k < 6; k++) geotransform6[k] = a.geotransform.size() == 6 ? a.geotransform[k] : 0.0;
  if (projection && projection_capacity > 0) {
    std::strncpy(projection, a.projection.c_str(), projection_capacity - 1);
    projection[projection_capacity - 1] = 0;
  }
  if (data) std::memcpy(data, a.data(), sizeof(float) * (size_t)a.width() * a.height());
  return 0;
}

// flats/flat_resolution.hpp:588-607: barnes_flat_resolution_d8(elevations, flowdirs, alter = false) -- what
// apps/rd_d8_flowdirs.cpp:18 ships: d8_flow_directions, resolve_flats_barnes (:448-515), d8_flow_flats (:97-116)
void ref_barnes_flat_resolution_d8_f32(const float *dem, int w, int h, float nodata, uint8_t *dirs, int32_t *mask_out,
                                       int32_t *labels_out) {
  Array2D<float> a(w, h);
  std::memcpy(a.data(), dem, sizeof(float) * (size_t)w * h);
  a.setNoData(nodata);
  Array2D<uint8_t> d;
  d8_flow_directions(a, d);
  Array2D<int32_t> m, l;
  resolve_flats_barnes(a, d, m, l);
  if (mask_out) std::memcpy(mask_out, m.data(), sizeof(int32_t) * (size_t)w * h);
  if (labels_out) std::memcpy(labels_out, l.data(), sizeof(int32_t) * (size_t)w * h);
  d8_flow_flats(m, l, d);
  std::memcpy(dirs, d.data(), (size_t)w * h);
}

}  // extern "C"
