/*
 * richdem_b200.h -- C ABI of librichdem_b200.so (B200 / sm_100a).
 *
 * This is the drop-in boundary for RichDEM's depression-filling / flat-resolution /
 * flow-routing hot path.  Every entry point below replaces one reference function template
 * (cited as file:line relative to the RichDEM source tree) for the dtypes the Python API
 * uses on this path: elevations float32, accumulation float64, proportions float32 x 9,
 * direction grids uint8.  Buffers are plain row-major rasters, i = y*width + x
 * (reference include/richdem/common/Array2D.hpp:592-595) -- i.e. exactly what
 * richdem::Array2D<T>::data() or a C-contiguous numpy array hands over.
 *
 * Conventions
 *   - All functions return 0 on success, non-zero on failure; rdb200_last_error() then
 *     returns a static, thread-local, human-readable message (CUDA errors included).
 *     There is NO CPU fallback: without a usable sm_100 device every compute call fails.
 *   - "Host" entry points take host pointers, copy to the device, compute, and copy the
 *     result back before returning; nothing is retained after return (same ownership rules
 *     as the reference: caller-owned, mutated in place).
 *   - "dev" entry points take device pointers on the current device and run on the
 *     library's stream; they let callers chain stages without leaving HBM.
 *   - Calls are synchronous from the caller's point of view and must be made from one
 *     thread at a time (the reference holds the GIL for the whole call as well).
 *   - D8 neighbour numbering is the reference's (include/richdem/common/constants.hpp:44-45):
 *         2 3 4
 *         1 0 5
 *         8 7 6
 */
#ifndef RICHDEM_B200_H_
#define RICHDEM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDB200_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define RDB200_API __attribute__((visibility("default")))
#else
#define RDB200_API
#endif

/* ---- library lifetime ------------------------------------------------------------- */

/* Selects CUDA device `device` (>=0) for this process, creates the stream and workspace.
 * Idempotent for the same device.  Compute calls lazily do rdb200_init(current device). */
RDB200_API int rdb200_init(int device);
/* Releases the workspace, stream and cached descriptors. */
RDB200_API void rdb200_shutdown(void);
RDB200_API const char *rdb200_last_error(void);
RDB200_API int rdb200_version(void);

/* Run all subsequent work on the caller's CUDA stream (a cudaStream_t; NULL restores the
 * library's own stream).  Lets a caller bracket calls with its own CUDA events. */
RDB200_API int rdb200_set_stream(void *cuda_stream);

/* Counters of the most recent call (for benchmarks / roofline accounting). */
typedef struct rdb200_stats {
  int64_t cells;             /* width*height of the raster processed                       */
  int64_t kernel_launches;   /* CUDA kernels launched by the call                          */
  int64_t fill_rounds;       /* fill: global sweep rounds (one persistent launch each)     */
  int64_t fill_tile_visits;  /* fill: tiles loaded+relaxed, summed over rounds             */
  int64_t fill_tile_cells;   /* fill: cells per tile (visits*tile_cells = cells swept)     */
  int64_t fill_tile_iters;   /* fill: in-shared-memory relaxation passes, summed           */
  int64_t accum_rounds;      /* accumulation: frontier rounds                              */
  int64_t flat_bfs_levels;   /* flats: BFS levels (away + towards)                         */
  int64_t flat_cells_raised; /* flats: cells whose elevation changed                       */
  double ms_total;           /* device time of the whole call (CUDA events)                */
  double ms_main_kernel;     /* device time summed over launches of the dominant kernel    */
  double ms_h2d, ms_d2h;     /* host entry points only                                     */
} rdb200_stats;
RDB200_API int rdb200_get_stats(rdb200_stats *out);

/* Tunables (algorithm switches and debug aids; see struct Params in csrc/common.cuh for the list and the
 * defaults).  "reset_defaults" restores every one of them; "trim_workspace" returns the cached device scratch (kept
 * between calls so that repeated calls do not pay cudaMalloc) to the driver.  They are process-wide and survive
 * rdb200_shutdown / re-init.  The library is single-threaded by contract (see above). */
RDB200_API int rdb200_set_param(const char *name, int64_t value);

/* ---- host entry points: the reference functions they replace ----------------------- */

/* richdem::FillDepressions<Topology::D8>(Array2D<float>&)
 *   include/richdem/depressions/depressions.hpp:13-21 -> PriorityFlood_Zhou2016,
 *   include/richdem/depressions/Zhou2016.hpp:125-191 (pyrichdem: rdFillDepressionsD8,
 *   wrappers/pyrichdem/src/pywrapper.hpp:32).  In place.  NoData is not special (as in the
 *   reference).  Result is bit-identical to the reference. */
RDB200_API int rdb200_fill_depressions_d8_f32(float *dem, int32_t width, int32_t height);

/* richdem::FillDepressions<Topology::D4>(Array2D<float>&)
 *   include/richdem/depressions/depressions.hpp:16-17 -> PriorityFlood_Barnes2014<Topology::D4>,
 *   include/richdem/depressions/Barnes2014.hpp:230-304 (pyrichdem: rdFillDepressionsD4, pywrapper.hpp:33).
 *   Same engine with the 4-neighbour stencil; bit-identical. */
RDB200_API int rdb200_fill_depressions_d4_f32(float *dem, int32_t width, int32_t height);

/* richdem::ResolveFlatsEpsilon(Array2D<float>&)
 *   include/richdem/flats/flats.hpp:21-28 (GetFlatMask + ResolveFlatsEpsilon_Barnes2014,
 *   include/richdem/flats/Barnes2014.hpp:398-467, 496-550); pyrichdem rdResolveFlatsEpsilon
 *   (pywrapper.hpp:37).  In place; bit-identical. */
RDB200_API int rdb200_resolve_flats_epsilon_f32(float *dem, int32_t width, int32_t height, float nodata);

/* richdem::GetFlatMask (include/richdem/flats/Barnes2014.hpp:398-467): the int32 increment
 * mask and a flat id per cell (0 = not in a drainable flat; ids are arbitrary but equal
 * within one flat -- the reference's ids are traversal-order dependent too). */
RDB200_API int rdb200_get_flat_mask_f32(const float *dem, int32_t *flat_mask, int32_t *labels, int32_t width,
                             int32_t height, float nodata);

/* richdem::d8_flow_directions(const Array2D<float>&, Array2D<uint8_t>&)
 *   include/richdem/flowmet/d8_flowdirs.hpp:96-123 (+ d8_FlowDir :32-74).  Codes 0..8,
 *   255 = NoData (include/richdem/common/constants.hpp:76).  Bit-identical. */
RDB200_API int rdb200_d8_flow_directions_f32(const float *dem, uint8_t *flowdirs, int32_t width,
                                  int32_t height, float nodata);

/* richdem::barnes_flat_resolution_d8(Array2D<float>&, Array2D<uint8_t>&, bool alter)
 *   include/richdem/flats/flat_resolution.hpp:588-607 (apps/rd_d8_flowdirs.cpp:18 ships it with alter = false):
 *   d8_flow_directions, then the increment mask / labels of the flats the DIRECTION grid shows (resolve_flats_barnes,
 *   :448-515) and either directions inside the flats from the mask (d8_flow_flats, :97-116; dem untouched) or, with
 *   alter != 0, the elevations altered by the mask (d8_flats_alter_dem, :540-586; dem updated in place) and the
 *   directions recomputed.  Bit-identical. */
RDB200_API int rdb200_d8_flow_directions_flats_f32(float *dem, uint8_t *flowdirs, int32_t width, int32_t height,
                                                   float nodata, int32_t alter);

/* richdem::d8_flow_accum(const Array2D<uint8_t>&, Array2D<int32_t>&)
 *   include/richdem/methods/d8_methods.hpp:47-139.  NoData direction = 255 -> area -1. */
RDB200_API int rdb200_d8_flow_accum_u8_i32(const uint8_t *flowdirs, int32_t *area, int32_t width,
                                int32_t height);

/* richdem::FM_D8 / FM_OCallaghan<D8>(const Array2D<float>&, Array3D<float>&)
 *   include/richdem/flowmet/OCallaghan1984.hpp:13-77,81-84.  props9 is [y][x][9] float
 *   (include/richdem/common/Array3D.hpp:203-206).  Bit-identical. */
RDB200_API int rdb200_fm_d8_f32(const float *dem, float *props9, int32_t width, int32_t height, float nodata);

/* richdem::FM_Tarboton / FM_Dinfinity  include/richdem/flowmet/Tarboton1997.hpp:14-149.
 * Facet choice identical; proportions within 1 float ulp (device atan2 vs libm). */
RDB200_API int rdb200_fm_tarboton_f32(const float *dem, float *props9, int32_t width, int32_t height,
                           float nodata);

/* richdem::FM_D4 = FM_OCallaghan<Topology::D4>  include/richdem/flowmet/OCallaghan1984.hpp:13-77,89-91.
 * As in the reference, the single receiver's proportion 1 is stored in slot n of the D4 numbering
 * (1 = W, 2 = N, 3 = E, 4 = S; common/constants.hpp:53-54).  Bit-identical. */
RDB200_API int rdb200_fm_d4_f32(const float *dem, float *props9, int32_t width, int32_t height, float nodata);
/* richdem::FM_Quinn  include/richdem/flowmet/Quinn1991.hpp:12-16 (= FM_Holmgren with exponent 1); bit-identical.
 * richdem::FM_Holmgren(elevations, props, xparam)  include/richdem/flowmet/Holmgren1994.hpp:13-83.
 * richdem::FM_Freeman(elevations, props, xparam)   include/richdem/flowmet/Freeman1991.hpp:13-80.
 * Same mixed float/double arithmetic as the reference; with an exponent other than 1 the device pow differs from
 * libm by <= 2 ulp (double), i.e. proportions within 1 float ulp. */
RDB200_API int rdb200_fm_quinn_f32(const float *dem, float *props9, int32_t width, int32_t height, float nodata);
RDB200_API int rdb200_fm_holmgren_f32(const float *dem, float *props9, int32_t width, int32_t height, float nodata,
                                      double xparam);
RDB200_API int rdb200_fm_freeman_f32(const float *dem, float *props9, int32_t width, int32_t height, float nodata,
                                     double xparam);

/* richdem::FA_D4 / FA_Quinn / FA_Holmgren / FA_Freeman(const Array2D<float>&, Array2D<double>&[, xparam])
 *   include/richdem/methods/flow_accumulation.hpp:28,19,18,20 (pyrichdem: pywrapper.hpp:65,58,57,59).
 *   FM_x into a device-resident proportions array (36 B/cell, never leaves HBM) + the generic accumulation below.
 *   accum_inout arrives holding the weights. */
RDB200_API int rdb200_fa_d4_f32_f64(const float *dem, double *accum_inout, int32_t width, int32_t height, float nodata);
RDB200_API int rdb200_fa_quinn_f32_f64(const float *dem, double *accum_inout, int32_t width, int32_t height,
                                       float nodata);
RDB200_API int rdb200_fa_holmgren_f32_f64(const float *dem, double *accum_inout, int32_t width, int32_t height,
                                          float nodata, double xparam);
RDB200_API int rdb200_fa_freeman_f32_f64(const float *dem, double *accum_inout, int32_t width, int32_t height,
                                         float nodata, double xparam);

/* richdem::TA_slope_riserun / TA_slope_percentage / TA_slope_degrees / TA_slope_radians / TA_aspect / TA_curvature /
 * TA_planform_curvature / TA_profile_curvature(const Array2D<float>&, Array2D<float>&, float zscale)
 *   include/richdem/methods/terrain_attributes.hpp:370-538 over TerrainProcessor (:336-354) and the per-cell
 *   formulas (:154-322); pyrichdem TerrainAttribute (wrappers/pyrichdem/richdem/__init__.py:735-794).
 *   NoData cells of the input (== nodata_in) become nodata_out (the output raster's own NoData, -9999 from
 *   pyrichdem); neighbours that are NoData or outside the raster take the centre's value.  cell_x / cell_y are
 *   |geotransform[1]| and |geotransform[5]| (Array2D.hpp:1387-1399).  Double arithmetic in the reference's order
 *   without fused multiply-adds: slopes (rise/run, percentage) and the three curvatures are bit-identical to a
 *   stock x86-64 build of the reference; the attributes through atan / atan2 (degrees, radians, aspect) are within
 *   1 float ulp (device libm differs from glibc by <= 2 ulp of the double). */
enum {
  RDB200_TA_SLOPE_RISERUN = 0,
  RDB200_TA_SLOPE_PERCENTAGE = 1,
  RDB200_TA_SLOPE_DEGREES = 2,
  RDB200_TA_SLOPE_RADIANS = 3,
  RDB200_TA_ASPECT = 4,
  RDB200_TA_CURVATURE = 5,
  RDB200_TA_PLANFORM_CURVATURE = 6,
  RDB200_TA_PROFILE_CURVATURE = 7
};
RDB200_API int rdb200_terrain_attribute_f32(int32_t attribute, const float *dem, float *out, int32_t width, int32_t height,
                                            float nodata_in, float nodata_out, float zscale, double cell_x, double cell_y);

/* richdem::FlowAccumulation(const Array3D<float>&, Array2D<double>&)
 *   include/richdem/methods/flow_accumulation_generic.hpp:33-100 (pyrichdem
 *   "FlowAccumulation", wrappers/pyrichdem/src/pywrapper.cpp:50).  accum arrives holding the
 *   per-cell weights and leaves holding the accumulation; NoData cells (props slot 0 == -2)
 *   become -1.  Flow out of raster-edge cells is ignored (the reference FM_* never emit it). */
RDB200_API int rdb200_flow_accumulation_props_f64(const float *props9, double *accum_inout, int32_t width,
                                       int32_t height);

/* richdem::FA_D8 / FA_Tarboton(=FA_Dinfinity)(const Array2D<float>&, Array2D<double>&)
 *   include/richdem/methods/flow_accumulation.hpp:27,16,17 (pywrapper.hpp:64,55,56).
 *   Fused: the 36 B/cell proportions array is never materialised.  accum_inout as above;
 *   pass accum_is_ones != 0 to promise that every weight is 1.0 (skips the upload). */
RDB200_API int rdb200_fa_d8_f32_f64(const float *dem, double *accum_inout, int32_t width, int32_t height,
                         float nodata, int32_t accum_is_ones);
RDB200_API int rdb200_fa_tarboton_f32_f64(const float *dem, double *accum_inout, int32_t width,
                               int32_t height, float nodata, int32_t accum_is_ones);

/* ---- device entry points (pointers into HBM of the current device) ----------------- */

RDB200_API int rdb200_dev_fill_depressions_d8_f32(float *d_dem, int32_t width, int32_t height);
RDB200_API int rdb200_dev_fill_depressions_d4_f32(float *d_dem, int32_t width, int32_t height);
RDB200_API int rdb200_dev_resolve_flats_epsilon_f32(float *d_dem, int32_t width, int32_t height, float nodata);
RDB200_API int rdb200_dev_d8_flow_directions_f32(const float *d_dem, uint8_t *d_flowdirs, int32_t width,
                                      int32_t height, float nodata);
RDB200_API int rdb200_dev_d8_flow_directions_flats_f32(float *d_dem, uint8_t *d_flowdirs, int32_t width, int32_t height,
                                                       float nodata, int32_t alter);
RDB200_API int rdb200_dev_d8_flow_accum_u8_i32(const uint8_t *d_flowdirs, int32_t *d_area, int32_t width,
                                    int32_t height);
RDB200_API int rdb200_dev_fm_d8_f32(const float *d_dem, float *d_props9, int32_t width, int32_t height,
                         float nodata);
RDB200_API int rdb200_dev_fm_tarboton_f32(const float *d_dem, float *d_props9, int32_t width, int32_t height,
                               float nodata);
/* method: 0 FM_D8, 1 FM_Tarboton, 2 FM_D4, 3 FM_Holmgren (xparam; FM_Quinn = 1.0), 4 FM_Freeman (xparam) */
RDB200_API int rdb200_dev_fm_method_f32(int32_t method, const float *d_dem, float *d_props9, int32_t width,
                                        int32_t height, float nodata, double xparam);
RDB200_API int rdb200_dev_fa_method_f32_f64(int32_t method, const float *d_dem, double *d_accum_inout, int32_t width,
                                            int32_t height, float nodata, double xparam);
RDB200_API int rdb200_dev_terrain_attribute_f32(int32_t attribute, const float *d_dem, float *d_out, int32_t width,
                                                int32_t height, float nodata_in, float nodata_out, float zscale,
                                                double cell_x, double cell_y);
RDB200_API int rdb200_dev_flow_accumulation_props_f64(const float *d_props9, double *d_accum_inout,
                                           int32_t width, int32_t height);
RDB200_API int rdb200_dev_fa_d8_f32_f64(const float *d_dem, double *d_accum_inout, int32_t width,
                             int32_t height, float nodata, int32_t accum_is_ones);
RDB200_API int rdb200_dev_fa_tarboton_f32_f64(const float *d_dem, double *d_accum_inout, int32_t width,
                                   int32_t height, float nodata, int32_t accum_is_ones);

/* Seeded synthetic fractal DEM (value-noise fBm, float32, no NaN / NoData) generated in HBM;
 * benchmark input only (the reference's generate_perlin_terrain is single-octave/double:
 * src/terrain_generation/terrain_generation.cpp:11-24).  Cell (x,y) of the band is global
 * cell (x, y0+y) of a raster `full_height` rows tall, so row bands of one raster agree. */
RDB200_API int rdb200_dev_generate_fbm_f32(float *d_dem, int32_t width, int32_t height, int32_t y0,
                                uint32_t seed, int32_t octaves, float quantum);

/* ---- multi-GPU: one process per GPU, the raster cut into row bands ----------------------------------------------
 * (the reference's own distributed path is an MPI tile farm: programs/parallel_priority_flood/main.cpp:394-548.)
 * Rank r of `world` holds ghost_top + owned + ghost_bottom rows in HBM (ghost_top = r > 0, ghost_bottom = r < world-1;
 * the ghost rows of the elevation raster hold the neighbouring bands' edge rows).  The rdb200_mgpu_* calls are
 * collective: every rank calls them in the same order.  A communicator is either NCCL (the product path; libnccl.so.2
 * is opened at run time) or a pair of caller-supplied callbacks (how the CPU tests run the same C++ drivers over gloo).
 *
 *   rank 0:  rdb200_nccl_unique_id(id)   ... ship the 128 bytes to every rank (MPI_Bcast, a file, torch.distributed) ...
 *   all:     rdb200_init(local_gpu); rdb200_comm_create_nccl(&comm, rank, world, id);
 *            rdb200_mgpu_fill_depressions_d8_f32(comm, d_band, W, rows, gt, gb, row0, H, NULL);
 *            rdb200_mgpu_fa_f32_f64(comm, d_band, d_accum, W, rows, nodata, gt, gb, 0, 1, NULL);
 */
typedef struct rdb200_comm rdb200_comm;
enum { RDB200_MAX_F32 = 0, RDB200_MIN_F32 = 1, RDB200_MAX_I32 = 2, RDB200_SUM_I32 = 3 };
/* exchange `bytes` with rank-1 (send_up / recv_up) and rank+1 (send_dn / recv_dn); a side without a neighbour gets
 * null pointers.  Pointers are device pointers of the calling rank; the call returns when the data has arrived. */
typedef int (*rdb200_exchange_fn)(void *user, const void *send_up, void *recv_up, const void *send_dn, void *recv_dn,
                                  size_t bytes);
/* in-place all-reduce of `count` elements, op = one of RDB200_MAX_F32 ... */
typedef int (*rdb200_allreduce_fn)(void *user, void *buf, size_t count, int op);
RDB200_API int rdb200_nccl_unique_id(uint8_t *out128);
RDB200_API int rdb200_comm_create_nccl(rdb200_comm **comm, int32_t rank, int32_t world, const uint8_t *id128);
RDB200_API int rdb200_comm_create_callbacks(rdb200_comm **comm, int32_t rank, int32_t world, void *user,
                                            rdb200_exchange_fn exchange, rdb200_allreduce_fn allreduce);
RDB200_API int rdb200_comm_destroy(rdb200_comm *comm);

/* FillDepressions<D8> (include/richdem/depressions/depressions.hpp:13-21) over row bands, in place on `d_band`
 * (local_rows x width, ghost rows included; their contents are ignored on entry and hold the neighbours' filled edge
 * rows on return).  row0 = global row of local row 0, height = rows of the whole raster.  Same result as the single-GPU
 * call, bit for bit.  *exchange_rounds (optional): halo exchanges done. */
RDB200_API int rdb200_mgpu_fill_depressions_d8_f32(const rdb200_comm *comm, float *d_band, int32_t width, int32_t local_rows,
                                                   int32_t ghost_top, int32_t ghost_bottom, int32_t row0, int32_t height,
                                                   int32_t *exchange_rounds);
/* FA_D8 (dinf = 0) / FA_Tarboton (dinf = 1) (include/richdem/methods/flow_accumulation.hpp:27,16) over row bands.
 * d_band_dem: elevations incl. ghost rows (neighbours' rows); d_band_accum_inout: weights in / accumulation out on the
 * owned rows (the ghost rows are scratch); accum_is_ones as in rdb200_fa_d8_f32_f64. */
RDB200_API int rdb200_mgpu_fa_f32_f64(const rdb200_comm *comm, const float *d_band_dem, double *d_band_accum_inout,
                                      int32_t width, int32_t local_rows, float nodata, int32_t ghost_top,
                                      int32_t ghost_bottom, int32_t dinf, int32_t accum_is_ones, int32_t *exchange_rounds);

/* ---- row-band (multi-GPU) fill: one band per GPU, halo rows exchanged by the caller -- */
/* The band raster handed in is (band_rows + ghost rows) x width.  Its first and last rows are
 * boundary conditions that the solver never changes: a real raster border row, or a ghost
 * row holding the neighbouring band's current water level (+inf before the first exchange).
 * Protocol per GPU: begin -> { run ; read own edge rows ; exchange ; update ghost rows } until
 * no band changed anywhere -> finish.  See richdem_b200/sharded.py. */
typedef struct rdb200_fill_state rdb200_fill_state;
RDB200_API int rdb200_dev_fill_begin(rdb200_fill_state **state, const float *d_dem, int32_t width,
                          int32_t height);
/* Multigrid start for a band (parameter fill_multigrid): `d_coarse` is the FILLED k x k max-pooled raster of the
 * WHOLE raster (coarse_width columns; every GPU solves that small raster itself), `row_offset` the global row of the
 * band's row 0.  Interior cells start at their block's coarse water level (an upper bound of the answer) instead of
 * +inf; the caller presets ghost rows the same way.  Everything else as rdb200_dev_fill_begin. */
RDB200_API int rdb200_dev_fill_begin_lifted(rdb200_fill_state **state, const float *d_dem, int32_t width, int32_t height,
                                 const float *d_coarse, int32_t coarse_width, int32_t pool, int32_t row_offset);
/* k x k max-pooling of rows [row_offset, row_offset + height) of a raster into the rows of the full coarse raster
 * (coarse_width x coarse_height, pre-filled by the caller, e.g. with -inf) that they touch; entries are combined
 * with max, so bands that share a coarse row can be merged with a MAX all-reduce. */
/* Multigrid V-cycle pieces for bands (parameter fill_vcycle):
 *   blockmax : k x k block maxima of the band's current water surface over its rows [skip_top, height - skip_bottom)
 *              (the owned rows), max-combined into the full coarse array (pre-filled with -inf; bands merge by MAX);
 *   relax_from: depression filling of `d_dem` started from the upper bound in `d_w_inout` (receives the result);
 *   prolong  : every interior cell of the band drops to its block's coarse level where that is lower; the tiles
 *              touched are queued for the next run; *tiles_lowered = how many. */
RDB200_API int rdb200_dev_fill_blockmax(rdb200_fill_state *state, float *d_blockmax, int32_t coarse_width, int32_t coarse_height,
                             int32_t pool, int32_t row_offset, int32_t skip_top, int32_t skip_bottom);
RDB200_API int rdb200_dev_fill_relax_from_f32(const float *d_dem, float *d_w_inout, int32_t width, int32_t height);
RDB200_API int rdb200_dev_fill_prolong(rdb200_fill_state *state, const float *d_coarse, int32_t coarse_width, int32_t pool,
                            int32_t row_offset, int32_t *tiles_lowered);
RDB200_API int rdb200_dev_maxpool_rows_f32(const float *d_src, int32_t width, int32_t height, int32_t row_offset, int32_t pool,
                                float *d_coarse, int32_t coarse_width, int32_t coarse_height);
/* Relax to the local fixed point (or for about `fill_band_rounds` sweep rounds when that parameter
 * is set).  *changed_rows: bit0 = row 1 changed, bit1 = row height-2 changed during this call (the
 * rows a neighbouring band holds as ghosts); bit2 = tiles are still active (call again). */
RDB200_API int rdb200_dev_fill_run(rdb200_fill_state *state, int32_t *changed_rows);
/* Copy water-level row y (0..height-1) to d_row[width]. */
RDB200_API int rdb200_dev_fill_read_row(rdb200_fill_state *state, int32_t y, float *d_row);
/* Replace boundary row y (0 or height-1) by d_row (element-wise <= the old values). */
RDB200_API int rdb200_dev_fill_update_row(rdb200_fill_state *state, int32_t y, const float *d_row);
/* Write the filled band to d_out (height x width, boundary rows included) and free state. */
RDB200_API int rdb200_dev_fill_finish(rdb200_fill_state *state, float *d_out);

/* ---- row-band (multi-GPU) flow accumulation ---------------------------------------------- */
/* Same role as FA_D8 / FA_Tarboton (include/richdem/methods/flow_accumulation.hpp:27,16) for one
 * row band.  The local raster (elevations and accumulation) is ghost_top + owned + ghost_bottom
 * rows; ghost elevation rows must hold the neighbouring bands' rows.  The ghost rows of the
 * accumulation array are scratch (parking slots for flow that leaves the band).
 * d_dem and d_accum_inout must stay valid until finish (the unit-weight D8 path computes the flow codes of the whole
 * band in its first run, after the neighbours' edge codes have arrived).
 * Protocol per GPU: begin -> exchange edge codes (get_edge_codes / set_ghost_codes) ->
 *   repeat { run ; take_outflow per side ; exchange ; apply_inflow per side } until no rank sent
 *   anything -> finish.  See richdem_b200/sharded.py. */
typedef struct rdb200_facc_state rdb200_facc_state;
RDB200_API int rdb200_dev_facc_begin(rdb200_facc_state **state, const float *d_dem, double *d_accum_inout,
                                     int32_t width, int32_t height, float nodata, int32_t ghost_top,
                                     int32_t ghost_bottom, int32_t dinf, int32_t accum_is_ones);
/* which: 0 = top side, 1 = bottom side.  Flow codes (1 byte/cell, + float rmax for D-infinity) of
 * my first/last owned row, to be installed as the neighbour's ghost codes. */
RDB200_API int rdb200_dev_facc_get_edge_codes(rdb200_facc_state *state, int32_t which, uint8_t *d_code_row,
                                              float *d_rmax_row);
RDB200_API int rdb200_dev_facc_set_ghost_codes(rdb200_facc_state *state, int32_t which,
                                               const uint8_t *d_code_row, const float *d_rmax_row);
/* Walk from the current frontier (first call: from all sources).  sent_*: number of flow parcels
 * parked in the top / bottom ghost row by this run. */
RDB200_API int rdb200_dev_facc_run(rdb200_facc_state *state, int32_t *sent_top, int32_t *sent_bottom);
/* Move the parked flow of one ghost row out (sum per cell, number of parcels per cell) and clear it. */
RDB200_API int rdb200_dev_facc_take_outflow(rdb200_facc_state *state, int32_t which, double *d_sum_row,
                                            int32_t *d_cnt_row);
/* Add a neighbour's outflow to my first/last owned row and release the cells it completes. */
RDB200_API int rdb200_dev_facc_apply_inflow(rdb200_facc_state *state, int32_t which, const double *d_sum_row,
                                            const int32_t *d_cnt_row);
RDB200_API int rdb200_dev_facc_finish(rdb200_facc_state *state);

/* ---- row-band (multi-GPU) flat resolution -------------------------------------------------- */
/* Same role as ResolveFlatsEpsilon (include/richdem/flats/flats.hpp:21-28) for one row band; the
 * local elevation raster (ghost_top + owned + ghost_bottom rows, ghost rows = the neighbours' rows)
 * is modified in place on the owned rows.  The steps run the single-GPU kernels on the local
 * raster; between them the caller moves the seam rows (see the protocol in csrc/flats.cu and
 * richdem_b200/sharded.py: resolve_flats_band). */
typedef struct rdb200_flats_state rdb200_flats_state;
RDB200_API int rdb200_dev_flats_begin(rdb200_flats_state **state, float *d_dem, int32_t width, int32_t height,
                                      float nodata, int32_t ghost_top, int32_t ghost_bottom);
RDB200_API int rdb200_dev_flats_arrays(rdb200_flats_state *state, uint64_t *out6);
RDB200_API int rdb200_dev_flats_edges(rdb200_flats_state *state);
RDB200_API int rdb200_dev_flats_components(rdb200_flats_state *state);
RDB200_API int rdb200_dev_flats_labels(rdb200_flats_state *state);
/* The distance state returned here speaks the row-band fill protocol: rdb200_dev_fill_run /
 * _read_row / _update_row; hand it back to gradient_end (do not call rdb200_dev_fill_finish). */
RDB200_API int rdb200_dev_flats_gradient_begin(rdb200_flats_state *state, int32_t away,
                                               rdb200_fill_state **dist_state);
RDB200_API int rdb200_dev_flats_gradient_end(rdb200_flats_state *state, int32_t away,
                                             rdb200_fill_state *dist_state);
RDB200_API int rdb200_dev_flats_apply(rdb200_flats_state *state);
RDB200_API int rdb200_dev_flats_finish(rdb200_flats_state *state);

#ifdef __cplusplus
}
#endif
#endif /* RICHDEM_B200_H_ */
