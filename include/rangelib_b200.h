/* rangelib_b200 -- C ABI of the B200-native batched 2-D ray casting path.
 *
 * This is the drop-in boundary: the entry points below are exactly what the reference's
 * Cython binding (range_libc, /root/reference/pywrapper/RangeLibc.pyx) needs from a backend
 * for its RangeMethod classes.  Each declaration cites the reference interface it replaces.
 * Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *  - every function returns 0 on success and a negative RL_E_* code on failure;
 *    rl_last_error() returns a thread-local message for the last failure.  (The reference
 *    returns void and prints / throws std::string / is UB; see INTEGRATION.md.)
 *  - data pointers (ins, angles, obs, ranges, outs, weights) may be HOST or DEVICE memory;
 *    the library detects which with cudaPointerGetAttributes.
 *      host   : inputs are staged to the device, the call returns after the results are back
 *               in the caller's buffer (same blocking semantics as the reference);
 *      device : the kernels run in place on the handle's stream and the call returns without
 *               synchronising (rl_method_synchronize() or the caller's own stream sync).
 *    All data pointers of one call must live on the same side.
 *  - buffers are caller-owned and borrowed for the duration of the call, never retained
 *    (RangeLibc.pyx:208-225 passes numpy buffers the same way).
 *  - handles are thread-compatible: use one rl_method per host thread / stream.
 *  - occupancy grids are x-major bytes: occ[x*H + y] != 0  <=>  OMap::grid[x][y]
 *    (/root/reference/includes/RangeLib.h:126).
 */
#ifndef RANGELIB_B200_H
#define RANGELIB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rl_map rl_map;
typedef struct rl_method rl_method;

/* range method kinds (RangeLib.h: BresenhamsLine :691, RayMarching :922 / RayMarchingGPU :774,
 * CDDTCast :971; PCDDT = CDDTCast + prune() :1176; GiantLUTCast :1772) */
enum { RL_BL = 0, RL_RM = 1, RL_CDDT = 2, RL_PCDDT = 3, RL_GLT = 4 };

enum {
  RL_OK = 0,
  RL_E_INVALID = -1,   /* bad argument */
  RL_E_CUDA = -2,      /* CUDA runtime failure (message has the cudaError string) */
  RL_E_NO_DEVICE = -3, /* no usable sm_100 device: the library has NO CPU fallback */
  RL_E_STATE = -4,     /* call not valid for this handle (e.g. prune on RM, no sensor model) */
  RL_E_MIXED = -5      /* host and device pointers mixed in one call */
};

const char* rl_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
uint64_t rl_stat_kernel_launches(void);

/* ---- OMap (RangeLib.h:121-322; Python PyOMap RangeLibc.pyx:130-198) ---------------------- */
/* replaces OMap(w,h)+grid fill (RangeLibc.pyx:138-145); copies occ */
int rl_map_create(const uint8_t* occ_xmajor, int width, int height, rl_map** out);
/* replaces the world_* field stores at RangeLibc.pyx:160-166 / 174-180; defaults 1,0,0,0,0,1 */
int rl_map_set_world(rl_map* map, float scale, float angle, float origin_x, float origin_y, float sin_angle,
                     float cos_angle);
int rl_map_width(const rl_map* map);
int rl_map_height(const rl_map* map);
/* OMap::get / isOccupied (RangeLib.h:203-210): 1 occupied, 0 free or out of bounds */
int rl_map_is_occupied(const rl_map* map, int x, int y);
/* copy the grid out (x-major bytes) */
int rl_map_get(const rl_map* map, uint8_t* out_xmajor);
/* extension for dynamic maps (BASELINE config 4): overwrite the w*h patch at (x0,y0);
 * patch is x-major patch[(x-x0)*h + (y-y0)].  Host side only; see rl_method_update_map. */
int rl_map_update(rl_map* map, const uint8_t* patch_xmajor, int x0, int y0, int w, int h);
void rl_map_destroy(rl_map* map);

/* ---- RangeMethod construction ------------------------------------------------------------ */
/* replaces BresenhamsLine(OMap,mr) :694, RayMarching(OMap,mr) :925, RayMarchingGPU(OMap,mr) :777,
 * CDDTCast(OMap,mr,td) :974 (+ prune for RL_PCDDT), GiantLUTCast(OMap,mr,td) :1781 (the W*H*td uint16
 * table is filled on the device by the RM kernel).  The map (and its world parameters) is
 * copied, like the reference's by-value OMap.  All acceleration structures (distance
 * transform, CDDT tables) are BUILT ON THE DEVICE and stay resident there.
 * device < 0 selects the current CUDA device. */
int rl_method_create(int kind, const rl_map* map, float max_range, unsigned theta_discretization, int device,
                     rl_method** out);
void rl_method_destroy(rl_method* m);
/* CDDTCast::prune(max_range) RangeLib.h:1176 (PyCDDTCast.prune RangeLibc.pyx:263-267) */
int rl_method_prune(rl_method* m, float max_range);
/* Binary checkpoint of a built (and possibly pruned) CDDT table.  The reference can only dump the table as YAML /
 * JSON text for its viewer (CDDTCast::serializeYaml / serializeJson RangeLib.h:1652-1735, main.cpp --cddt_save_path)
 * and cannot load it back; this stores the same content (theta_discretization, lut_translations, max_range, map size,
 * the sorted zero points of every bin) as the CSR arrays the kernels read.  rl_method_create_from_cddt builds a
 * CDDT / PCDDT handle for `map` from such a file WITHOUT rebuilding or pruning; the file must have been written for
 * the same occupancy grid (size and content hash are checked; RL_E_INVALID otherwise).  The map's world parameters
 * are taken from `map` as in rl_method_create. */
int rl_method_save_cddt(rl_method* m, const char* path);
/* max_range / theta_discretization of a handle and whether its table has been pruned (any pointer may be NULL) */
int rl_method_get_params(const rl_method* m, float* max_range, unsigned* theta_discretization, int* pruned);
int rl_method_create_from_cddt(const rl_map* map, const char* path, int device, rl_method** out);
/* run this handle's work on the given cudaStream_t (0 = CUDA's default stream).  A new handle
 * runs on a private non-blocking stream until this is called; rl_method_use_own_stream goes back. */
int rl_method_set_stream(rl_method* m, void* cuda_stream);
int rl_method_use_own_stream(rl_method* m);
int rl_method_synchronize(rl_method* m);
/* dynamic maps: apply an occupancy patch on the device and refresh the structures that depend
 * on it (BL: bit grid only; RM: distance transform rebuilt; CDDT: table rebuilt).
 * patch may be host or device memory. */
int rl_method_update_map(rl_method* m, const uint8_t* patch_xmajor, int x0, int y0, int w, int h);
/* the same for n patches in one call (two launches: all cells, then the bit tiles they touch): rects = n x
 * (x0, y0, w, h) (HOST ints); the patches' bytes are concatenated in `patches` (HOST or DEVICE), each x-major
 * inside its rectangle.  Patches must not overlap cell-wise (checked for n <= 4096; RL_E_INVALID); they need not
 * be 8-aligned and may share a bit tile. */
int rl_method_update_map_batch(rl_method* m, const uint8_t* patches_xmajor, const int* rects, int n);
/* Whole-map ingest on the device: replace the occupancy of an existing handle (same size) from a source image in
 * HOST or DEVICE memory and refresh the kind's structures -- a mapping pipeline that keeps its grid in HBM never
 * copies it to the host.
 *  - occupancy_grid: ROS nav_msgs/OccupancyGrid `data` (int8 row-major [rows][cols]) with the reference's
 *    PyOMap(OccupancyGrid) rule (RangeLibc.pyx:146-157): OMap(rows, cols), grid[x][y] = data[x*cols+y] > 10,
 *    so rows must equal the map width and cols its height;
 *  - rgba: RGBA8 rows as lodepng_decode32 yields them, with OMap(filename, threshold)'s conversion
 *    (RangeLib.h:189-199, RangeUtils.h:30-32): gray from bytes (2,1,0), occupied iff (int)gray < threshold;
 *    img_w / img_h must equal the map width / height. */
int rl_method_set_map_occupancy_grid(rl_method* m, const int8_t* data, int rows, int cols);
int rl_method_set_map_rgba(rl_method* m, const uint8_t* rgba, int img_w, int img_h, float threshold);
/* bytes of device memory held by the acceleration structure (RangeMethod::memory()) */
int64_t rl_method_memory(const rl_method* m);

/* ---- queries -------------------------------------------------------------------------------- */
/* RangeMethod::calc_range(x,y,heading) RangeLib.h:418 -- one ray, grid coordinates.  Works for
 * every kind including RM-on-GPU (the reference's RayMarchingGPU::calc_range only prints,
 * RangeLib.h:800-807). */
int rl_calc_range(rl_method* m, float x, float y, float heading, float* out);
/* RayMarchingGPU::calc_range_many(ins,outs,n) RangeLib.h:819-831: n rays (x,y,theta) AoS in GRID
 * coordinates, no world conversion. */
int rl_calc_range_many(rl_method* m, const float* ins, float* outs, int num_casts);
/* RangeMethod::numpy_calc_range(ins,outs,n) RangeLib.h:439-480 (Python calc_range_many):
 * WORLD coordinates, ROS conversion incl. the x/y swap, result scaled by world_scale. */
int rl_numpy_calc_range(rl_method* m, const float* ins, float* outs, int num_casts);
/* RangeMethod::numpy_calc_range_angles RangeLib.h:482-520 (Python calc_range_repeat_angles):
 * outs[i*num_angles + a] for particle i and angle a. */
int rl_numpy_calc_range_angles(rl_method* m, const float* ins, const float* angles, float* outs, int num_particles,
                               int num_angles);
/* RangeMethod::set_sensor_model(table,k) RangeLib.h:523-532.  REPLACES the table (the reference
 * appends rows when called twice; calling once is identical). table is HOST or DEVICE, k*k doubles. */
int rl_set_sensor_model(rl_method* m, const double* table, int table_width);
/* RangeMethod::eval_sensor_model(obs,ranges,outs,rays_per_particle,particles) RangeLib.h:533-555 */
int rl_eval_sensor_model(rl_method* m, const float* obs, const float* ranges, double* outs, int rays_per_particle,
                         int particles);
/* RangeMethod::calc_range_repeat_angles_eval_sensor_model RangeLib.h:558-612 -- fused: ranges
 * never leave the SM. */
int rl_calc_range_repeat_angles_eval_sensor_model(rl_method* m, const float* ins, const float* angles,
                                                  const float* obs, double* weights, int num_particles,
                                                  int num_angles);
/* RangeMethod::calc_range_many_radial_optimized(ins,outs,num_particles,num_rays,min_angle,max_angle)
 * RangeLib.h:616-676 (PyCDDTCast.calc_range_many_radial_optimized, RangeLibc.pyx:274-276): a lidar fan of
 * num_rays beams from min_angle to max_angle per pose; beam a <= num_rays/3 and the beam pi further round come
 * from one CDDTCast::calc_range_pair (:1521-1649), the beams between from calc_range.  outs is
 * [num_particles][num_rays]; as in the reference, beams beyond the last paired one are NOT written (outs is
 * read-modify-write for host pointers), and kinds without calc_range_pair store -world_scale in the paired
 * slots (:419).  Second beams that fall outside the row are dropped (the reference writes past it). */
int rl_calc_range_many_radial_optimized(rl_method* m, const float* ins, float* outs, int num_particles, int num_rays,
                                        float min_angle, float max_angle);

/* Multi-GPU form of the fused call (one process per GPU, particles sharded across ranks, map and
 * tables replicated): the kernel's epilogue stores this rank's `num_particles` weights directly into
 * the gathered weight array of EVERY rank -- peer_weights[r] is rank r's array as a peer-mapped
 * device pointer (e.g. torch symmetric memory / cudaIpc / cuMem fabric handle), written at
 * [offset, offset + num_particles).  The compute and the all-gather are one kernel; the caller only
 * needs a cross-rank barrier before reading.  Device pointers only; asynchronous on the handle's
 * stream.  (No counterpart in the reference, which is single-GPU.) */
int rl_calc_range_repeat_angles_eval_sensor_model_peers(rl_method* m, const float* ins, const float* angles,
                                                        const float* obs, double* const* peer_weights, int n_peers,
                                                        int64_t offset, int num_particles, int num_angles);

/* Signalled form: the same fused kernel also carries the synchronisation, so a step is ONE kernel per
 * rank and nothing else (no barrier launch, no NCCL call).
 *   rl_method_peers_init   weights0[r] / weights1[r]: rank r's two gathered weight arrays (double
 *                          buffering), flags[r]: rank r's int64[n_peers] flag array (zero-initialised),
 *                          all peer mapped; rank = this process.
 *   ..._signalled          launch epoch e (1, 2, ... counted per handle; every rank must issue the same
 *                          sequence): waits in-kernel until every rank has finished epoch e-1, stores this
 *                          rank's weights into buffer e & 1 of every rank at `offset`, and the last CTA
 *                          publishes flags[r][rank] = e everywhere.  *buffer_index = e & 1.
 *   rl_method_peers_wait   enqueue a wait (on the handle's stream) until every rank's slice of the
 *                          latest epoch has arrived; after it the gathered array may be read. */
int rl_method_peers_init(rl_method* m, double* const* weights0, double* const* weights1, int64_t* const* flags,
                         int n_peers, int rank);
int rl_calc_range_repeat_angles_eval_sensor_model_signalled(rl_method* m, const float* ins, const float* angles,
                                                            const float* obs, int64_t offset, int num_particles,
                                                            int num_angles, int* buffer_index);
int rl_method_peers_wait(rl_method* m);
/* The sharded particle-filter update as ONE blocking call with HOST buffers (what a multi-process particle filter
 * written against the reference's calc_range_repeat_angles_eval_sensor_model would call on each rank; replaces the
 * loop RangeLib.h:558-612 over ALL particles): this rank's `num_particles` poses (the slice starting at particle
 * `offset` of `n_total`) are copied to the device, the signalled update (one fused kernel, or cast + evaluation kernel for deep updates) computes their weights and stores them
 * into every rank's gathered array over NVLink, and after every rank's slice has arrived the whole gathered array
 * (n_total doubles) is copied into weights_all.  Needs rl_method_peers_init; every rank must call it the same number
 * of times.  With DEVICE pointers nothing blocks: the copy into weights_all is device-to-device on the handle's stream. */
int rl_calc_range_repeat_angles_eval_sensor_model_sharded(rl_method* m, const float* ins, const float* angles,
                                                          const float* obs, double* weights_all, int64_t offset,
                                                          int num_particles, int num_angles, int64_t n_total);

/* ---- particle-filter steps either side of the sensor update ------------------------
 * NOT part of the reference: range_libc ends at the weights (RangeLib.h:558-612); these are the steps its downstream
 * user (the mit-racecar particle_filter MCL loop named in the reference's README) runs on the host between two sensor
 * updates, offered here so that particles and weights can stay on the device.  The handle supplies device, stream and
 * scratch only (any kind).  Pointers: HOST (blocking) or DEVICE (asynchronous), not mixed.  Exact definitions:
 * range_libc_b200/csrc/rl_pf.cu; checker: oracle/pf_oracle.py. */
/* w_i <- pow(w_i, inv_squash) (skipped when inv_squash == 1), then w_i <- w_i / sum(w).  *sum_out (HOST, may be NULL;
 * non-NULL makes the call blocking) receives the sum of the squashed weights. */
int rl_pf_normalize_weights(rl_method* m, double* weights, int n, double inv_squash, double* sum_out);
/* systematic (low-variance) resampling with one uniform draw u0 in [0, 1): out_particles[j] = particles[i_j], i_j the
 * first i whose fixed-point (2^-40) cumulative weight exceeds ((u0 + j) / n) * total.  Order-independent, bit-exact. */
int rl_pf_resample(rl_method* m, const float* particles, const double* weights, float* out_particles, int n, double u0);
/* odometry step of n planar poses (x, y, theta), in place: x += cos(theta) dx - sin(theta) dy (+ noise[3i]),
 * y += sin(theta) dx + cos(theta) dy (+ noise[3i+1]), theta += dtheta (+ noise[3i+2]); noise: n x 3 floats or NULL. */
int rl_pf_motion_update(rl_method* m, float* particles, int n, float dx, float dy, float dtheta, const float* noise);

/* ---- table-level access for parity tests ---------------------------------------------------- */
/* the occupancy bytes resident on the device, x-major out[x*H+y] (after dynamic updates / device ingest). out: HOST */
int rl_debug_get_occ(rl_method* m, uint8_t* out);
/* distance transform, x-major out[x*H+y] (DistanceTransform::grid RangeLib.h:328); RL_RM only. out: HOST */
int rl_debug_get_dt(rl_method* m, float* out);
/* CDDT tables (CDDTCast::compressed_lut / lut_translations RangeLib.h:1746-1748) in CSR form.
 * widths/translations: theta_discretization entries each (may be NULL); returns the total
 * number of bins in *n_bins and of zero points in *n_values. */
int rl_debug_cddt_dims(rl_method* m, int64_t* n_bins, int64_t* n_values, int* widths, float* translations);
/* offsets: n_bins+1 int64, values: n_values floats; HOST buffers */
int rl_debug_cddt_dump(rl_method* m, int64_t* offsets, float* values);
/* tuning knob (RM, small launches): once a CTA (256 rays) has at most `rays` unfinished rays left they are
 * finished one per warp with all 32 lanes cooperating (default 16; 0 = never hand off).  Results are identical. */
int rl_debug_set_coop_threshold(rl_method* m, int rays);
/* tuning knob (RM): large batches use persistent warps with lane re-queuing (default 1) or the
 * one-ray-per-thread kernel (0).  Results are identical. */
int rl_debug_set_persistent(rl_method* m, int on);
/* tuning knob (fused call, clouds >= 32768 particles on structures larger than L2): process the particles in the
 * order of the 64x64-cell tile they stand in (default 1) or in caller order (0).  CDDT / PCDDT: 0 searches the zero
 * points directly even when the table is larger than L2 (default: through the L2-resident query index), 2 builds and
 * uses the index whatever the table size (tests).  Results are identical. */
int rl_debug_set_spatial_sort(rl_method* m, int on);
/* GiantLUTCast::giant_lut (RangeLib.h:1903) as out[(x*H + y)*td + i], W*H*td uint16; HOST buffer */
int rl_debug_glt_dump(rl_method* m, uint16_t* out);
/* device trig used by BL/RM (restated glibc sinf/cosf); HOST buffers; for tests */
int rl_debug_sincosf(const float* x, float* s, float* c, int n);

#ifdef __cplusplus
}
#endif
#endif /* RANGELIB_B200_H */
