/* lama_b200.h -- C-ABI of the B200-native LaMa hot path (liblama_b200.so).
 *
 * The reference (iris-ua/iris_lama) has no plugin / FFI layer: the particle-filter SLAM hot path sits
 * behind plain C++ classes.  Every entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  Plain pointers and sizes only; every function returns an
 * int status (0 = ok, < 0 = error, see LAMA_ERR_*) and lama_last_error() gives the message.
 * Handles are not thread safe: one host thread per handle (as the reference objects).
 *
 * Conventions
 *   poses       xyr[3] = (x, y, rotation) like lama::Pose2D(x, y, rotation), include/lama/pose2d.h:45
 *   SE2 states  state[4] = (cos, sin, tx, ty): the raw Sophus SE2 of Pose2D::state, pose2d.h:76
 *   scans       pts_xyz = N x 3 doubles (PointCloudXYZ::points), sensor_origin[3], sensor_quat_xyzw[4]
 *               (PointCloudXYZ::sensor_origin_ / sensor_orientation_), include/lama/types.h:111-120
 *   map cells   absolute unsigned map coordinates as produced by Map::w2m, include/lama/sdm/map.h:125
 */
#ifndef LAMA_B200_H
#define LAMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAMA_OK 0
#define LAMA_ERR_ARG (-1)
#define LAMA_ERR_CUDA (-2)
#define LAMA_ERR_NO_DEVICE (-3)
#define LAMA_ERR_WINDOW (-4)   /* the map grew outside the device directory window (raise dir_dim) */
#define LAMA_ERR_POOL (-5)     /* device patch pool exhausted (raise pool_slots) */
#define LAMA_ERR_OVERFLOW (-6) /* per-scan event log / brushfire heap overflow */
#define LAMA_ERR_STATE (-7)

/* message of the last failing call on this thread */
const char* lama_last_error(void);
/* library / build identification, e.g. "lama_b200 0.1 sm_100a" */
const char* lama_version(void);
/* number of visible CUDA devices (0 when none: every create call then fails with LAMA_ERR_NO_DEVICE) */
int lama_device_count(void);

/* device-side knobs shared by all front ends (no counterpart in the reference) */
typedef struct lama_device_options {
    int32_t device;      /* CUDA device ordinal */
    int32_t dir_dim;     /* map window = dir_dim x dir_dim patches of 32 x 32 cells, power of two, default 64 */
    int32_t pool_slots;  /* 4 KiB patches in the device pool, 0 = auto */
    int32_t max_beams;   /* largest scan accepted, default 2048 */
    int32_t timing;      /* 1: record CUDA-event times per kernel (lama_*_kernel_times) */
    uint64_t stream;     /* cudaStream_t to launch on (e.g. a torch stream), 0 = the handle creates its own */
} lama_device_options;

/* ------------------------------------------------------------------------------------------------
 * PFSlam2D -- include/lama/pf_slam2d.h:132-232, src/pf_slam2d.cpp:106-574
 * ------------------------------------------------------------------------------------------------ */
typedef struct lama_pf lama_pf;

typedef struct lama_pf_options { /* PFSlam2D::Options, pf_slam2d.h:132-185 */
    uint32_t particles;
    double srr, str, stt, srt;
    double meas_sigma, meas_sigma_gain;
    double trans_thresh, rot_thresh;
    double l2_max;
    double truncated_ray, truncated_range;
    double resolution;
    uint32_t patch_size; /* must be 32 */
    uint32_t max_iter;
    int32_t strategy;    /* 0 "gn", 1 "lm" (PFSlam2D::scanMatch always uses Gauss-Newton, pf_slam2d.cpp:423-427) */
    int32_t threads;     /* accepted for source compatibility; the particle loop runs on the GPU */
    uint32_t seed;       /* 0 = random_device, pf_slam2d.cpp:131-134 */
    /* particle sharding over GPUs (one process per GPU): this handle owns particles
       [shard_rank * particles / shard_count, (shard_rank + 1) * particles / shard_count) */
    uint32_t shard_rank, shard_count;
    lama_device_options dev;
} lama_pf_options;

/* fills the reference defaults (pf_slam2d.h:132-185); `particles` has none there and is set to 1 */
int lama_pf_options_default(lama_pf_options* o);
/* PFSlam2D::PFSlam2D(const Options&), pf_slam2d.cpp:106-138 */
int lama_pf_create(const lama_pf_options* o, lama_pf** out);
int lama_pf_destroy(lama_pf* h);
/* PFSlam2D::setPrior, pf_slam2d.cpp:146-149 */
int lama_pf_set_prior(lama_pf* h, const double xyr[3]);
/* bool PFSlam2D::update(surface, odometry, timestamp), pf_slam2d.cpp:178-312; *did_update = the bool */
int lama_pf_update(lama_pf* h, const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4],
                   const double odom_xyr[3], double timestamp, int* did_update);
/* Device-resident inputs: copy n_scans x n x 3 doubles into HBM once, then run update() on scan `index`
   without any host->device transfer of points (same semantics as lama_pf_update otherwise) */
int lama_pf_stage_scans(lama_pf* h, const double* pts_xyz, int n_scans, int n);
int lama_pf_update_staged(lama_pf* h, int index, const double sensor_origin[3], const double sensor_quat_xyzw[4], const double odom_xyr[3],
                          double timestamp, int* did_update);
/* bytes moved host->device [0] and device->host [1] by the update path since the last reset (reset != 0 also
   clears the kernel timers) */
int lama_pf_get_traffic(lama_pf* h, uint64_t bytes[2], int reset);
/* PFSlam2D::getPose (best particle), pf_slam2d.cpp:332-336 */
int lama_pf_get_pose(lama_pf* h, double xyr[3]);
int lama_pf_get_best_particle(lama_pf* h, int* idx);      /* getBestParticleIdx, pf_slam2d.cpp:314-330 */
int lama_pf_get_neff(lama_pf* h, double* neff);           /* getNeff, pf_slam2d.h:229 */
/* getParticles(): states P x 4, weights P x 3 = (weight, normalized_weight, weight_sum); either may be NULL */
int lama_pf_get_particles(lama_pf* h, double* states, double* weights);
/* Particle::poses history of one particle as xyr triples; returns the length in *count (cap = capacity) */
int lama_pf_get_trajectory(lama_pf* h, int particle, double* xyr, int cap, int* count);
/* indices drawn by the last PFSlam2D::resample (pf_slam2d.cpp:537-574); *count = 0 when the last update did not resample */
int lama_pf_get_last_resample(lama_pf* h, int32_t* idx, int* count);
/* The whole resampling history in 16 bytes: out[0] = resamplings so far (Summary::resample count, pf_slam2d.h:88-129), out[1] = FNV-1a
 * hash over (accepted-scan number, the P ancestor indices) of each of them, bytes little endian (src/pf_slam2d.cpp:537-553). */
int lama_pf_get_resample_digest(lama_pf* h, uint64_t out[2]);
/* PFSlam2D::Summary time buckets (include/lama/pf_slam2d.h:88-129, filled at src/pf_slam2d.cpp:251-311) as host wall-clock sums in ms:
 * {sampling (drawFromMotion), solve (scan matching: enqueue + wait), normalise, resample}; the map bucket is the device time of
 * lama_pf_kernel_times (the map update runs asynchronously behind the next scan's sampling). */
int lama_pf_get_summary(lama_pf* h, double ms[4]);
/* uint64_t PFSlam2D::getMemoryUsage() and getMemoryUsage(occmem, dmmem) (src/pf_slam2d.cpp:151-176) from Map::memory() (src/sdm/map.cpp:115-125: per patch
 * 72 bytes of table entry + cell bytes / use count of the shared patch): out = {total over the particles, occmem, dmmem}.  The two-argument overload of the
 * reference adds particle 0's maps P times; occmem / dmmem reproduce that.  On a sharded handle the sums run over this rank's particles.
 * The reference's distance map also owns a patch wherever an occupancy cell was touched; the device counts a distance patch as shared by no more
 * particles than the occupancy patch over the same cells (an upper estimate of the reference's bytes once particles diverge, exact while maps are
 * fully shared). */
int lama_pf_get_memory_usage(lama_pf* h, uint64_t out[3]);
/* const std::deque<double>& PFSlam2D::getTimestamps() (include/lama/pf_slam2d.h:205-206; only the first scan's stamp is ever pushed, src/pf_slam2d.cpp:187) */
int lama_pf_get_timestamps(lama_pf* h, double* stamps, int cap, int* count);
/* work counters of the last update and totals: {residual evals (as the reference would count), ray cells,
   distance-map pops, patches detached, GN iterations, resampled} */
int lama_pf_get_counters(lama_pf* h, uint64_t last[6], uint64_t total[6]);
/* accumulated CUDA-event kernel times in ms {match, raycast, brushfire, resample} and launches {same + misc} */
int lama_pf_kernel_times(lama_pf* h, double ms[4], uint64_t launches[5]);
/* Map::bounds of a particle's map (kind 0 occupancy, 1 distance): min/max cell, *patches = numOfPatches */
int lama_pf_map_bounds(lama_pf* h, int particle, int kind, uint32_t mn[2], uint32_t mx[2], int* patches);
/* Map::write (src/sdm/map.cpp:490-529) of getOccupancyMap(particle) (kind 0) / getDistanceMap(particle) (kind 1): the reference's
 * ".sdm" file -- IOHeader (map.h:95-103), DynamicDistanceMap's max_sqdist_ (dynamic_distance_map.cpp:200-203), then per patch
 * its id, its 1024 cells in the reference cell layout and the 128-byte known mask (container.cpp:143-163). */
int lama_pf_write_map(lama_pf* h, int particle, int kind, const char* path);
/* OccupancyMap::{getProbability, isFree, isOccupied, isUnknown}(Vector3ui) (include/lama/sdm/occupancy_map.h:57-76,
 * frequency_occupancy_map.cpp:110-172) of getOccupancyMap(particle) for n cells: prob[i] = getProbability, flags[i] bit 0 isFree,
 * bit 1 isOccupied, bit 2 isUnknown.  The Vector3d overloads are lama_w2m (Map::w2m, map.h:125-126) followed by this call. */
int lama_pf_occupancy_query(lama_pf* h, int particle, const uint32_t* cells_xy, int n, double* prob, uint8_t* flags);
/* getDistanceMap(particle)->distance(Vector3d, Vector3d* gradient) for n points (distance_map.h:66, dynamic_distance_map.cpp:66-92):
 * dist[n], grad[n][3] (grad may be NULL) */
int lama_pf_distance(lama_pf* h, int particle, const double* pts_xyz, int n, double* dist, double* grad);
int lama_w2m(double resolution, const double* pts_xyz, int n, uint32_t* cells_xy);
/* the grey image sdm::export_to_png encodes (src/sdm/export.cpp:46-96; PFSlam2D::saveOccImage / saveDistImage use it):
 * dims = {width, height} = the map's bounds in cells, pixel (u, v) at pixels[u + v * width]; pixels == NULL only sizes. */
int lama_pf_export_image(lama_pf* h, int particle, int kind, uint8_t* pixels, size_t cap, int dims[2]);
/* dense window of FrequencyOccupancyMap cells {occupied, visited} + Container "known" bit; arrays may be NULL */
int lama_pf_export_occupancy(lama_pf* h, int particle, uint32_t x0, uint32_t y0, int w, int hgt, uint16_t* occupied, uint16_t* visited,
                             uint8_t* known);
/* dense window of DynamicDistanceMap::distance_t fields (dynamic_distance_map.h:48-53) + known bit */
int lama_pf_export_distance(lama_pf* h, int particle, uint32_t x0, uint32_t y0, int w, int hgt, uint16_t* sqdist, uint8_t* valid,
                            uint8_t* known, int16_t* ox, int16_t* oy, uint8_t* queued);

/* ---- multi-GPU behind lama_pf_update: particles shard over shard_count ranks, one host thread / process per GPU.  Rank 0 asks for an
 * id (lama_shard_unique_id = ncclGetUniqueId), hands its 128 bytes to every rank by any means (MPI, a file, torch.distributed ...), and
 * every rank calls lama_pf_shard_connect() on its handle (created with shard_rank / shard_count and the SAME non-zero seed).  From then
 * on lama_pf_update() / lama_pf_update_staged() run the whole sharded step of PFSlam2D::update (src/pf_slam2d.cpp:178-312): match +
 * map update of the local particles (:254-266, :292-302), ONE NCCL all-gather of the match results, normalize / resample (:274-287,
 * :511-574) on identical bytes on every rank, NCCL send / recv of the maps of remote ancestors.  libnccl.so.2 is loaded at run time.
 * lama_pf_shard_stats: {collectives issued, bytes of maps received}. */
int lama_shard_unique_id(uint8_t id[128]);
int lama_pf_shard_connect(lama_pf* h, const uint8_t id[128]);
int lama_pf_shard_stats(lama_pf* h, uint64_t out[2]);

/* --- sharded (multi-GPU) operation: the caller moves the small per-scan vectors between ranks ------------
 * begin : predict (every rank draws the noise of ALL particles, keeping the RNG streams identical) + gate +
 *         scan matching of the local shard; local_out = P_local x 5 doubles (state[4], log-likelihood)
 * finish: takes the gathered P x 5 vector, normalises, decides on resampling and returns the systematic
 *         resampling indices (identical on every rank; rank 0's are broadcast for safety)
 * apply : applies the indices; new local particle k takes the maps of resident slot local_src[k], which is
 *         either a local particle (0 .. P_local-1) or a staging slot (P_local .. 2 P_local-1) previously
 *         filled with lama_pf_particle_unpack from a buffer packed on the ancestor's rank
 * map   : ray-cast + distance-map update of the local shard
 * *did_update of begin: 0 = gated (nothing to do), 1 = first scan handled completely, 2 = matched: the
 * caller must continue with finish [/ apply] / map_update                                                */
int lama_pf_shard_begin(lama_pf* h, const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4],
                        const double odom_xyr[3], double timestamp, int* did_update, double* local_out);
int lama_pf_shard_finish(lama_pf* h, const double* all_results, int* resampled, int32_t* idx);
int lama_pf_shard_apply(lama_pf* h, const int32_t* idx); /* single rank: ancestors are idx themselves */
int lama_pf_shard_apply_local(lama_pf* h, const int32_t* idx, const int32_t* local_src);
int lama_pf_shard_map_update(lama_pf* h);
/* serialise / restore the maps of one resident slot for migration between ranks (host buffers) */
int lama_pf_particle_pack_size(lama_pf* h, int local_particle, size_t* bytes);
int lama_pf_particle_pack(lama_pf* h, int local_particle, void* buf, size_t cap, size_t* used);
int lama_pf_particle_unpack(lama_pf* h, int local_particle, const void* buf, size_t bytes);

/* ------------------------------------------------------------------------------------------------
 * Slam2D -- include/lama/slam2d.h:91-161, src/slam2d.cpp:92-321
 * ------------------------------------------------------------------------------------------------ */
typedef struct lama_slam lama_slam;
typedef struct lama_slam_options { /* Slam2D::Options, slam2d.h:91-125 */
    double trans_thresh, rot_thresh, l2_max, truncated_ray, truncated_range, resolution;
    uint32_t patch_size, max_iter;
    int32_t strategy; /* 0 "gn", 1 "lm" (slam2d.cpp:226-233) */
    int32_t occupancy; /* 0 = FrequencyOccupancyMap as in Slam2D (slam2d.cpp:97); 1 = ProbabilisticOccupancyMap, the log-odds map of
                          src/sdm/probabilistic_occupancy_map.cpp (what LidarOdometry2D pairs with the same update loop) */
    int32_t transient_map;  /* Slam2D::Options::transient_map (slam2d.h:122): after every map update drop the patches that do not
                               meet the (doubled, pose-centred, 2 * maxDistance grown) AABB of the scan, slam2d.cpp:323-379 */
    int32_t lidar_odometry; /* run the handle as lama::LidarOdometry2D (src/lidar_odometry_2d.cpp:42-181): odom_xyr is ignored (may be
                               NULL), log-odds map, l2_max 1.0, rays keep their last metre, map updated after 0.1 m / 0.5 rad of
                               estimated motion, transient map always on; lama_slam_get_pose returns LidarOdometry2D::odom */
    lama_device_options dev;
} lama_slam_options;
int lama_slam_options_default(lama_slam_options* o);
int lama_slam_create(const lama_slam_options* o, lama_slam** out);
int lama_slam_destroy(lama_slam* h);
int lama_slam_set_pose(lama_slam* h, const double xyr[3]);                   /* Slam2D::setPose, slam2d.h:147 */
int lama_slam_update(lama_slam* h, const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4],
                     const double odom_xyr[3], double timestamp, int* did_update); /* Slam2D::update, slam2d.cpp:143-198 */
int lama_slam_get_pose(lama_slam* h, double xyr[3]);
int lama_slam_get_state(lama_slam* h, double state[4]);
int lama_slam_get_processed_cells(lama_slam* h, uint32_t* n);               /* getNumberOfProcessedCells, slam2d.h:139 */
int lama_slam_get_map_stats(lama_slam* h, uint64_t stats[2]);                /* {map updates so far, patches deleted by the transient map} */
int lama_slam_get_counters(lama_slam* h, uint64_t last[6], uint64_t total[6]);
int lama_slam_kernel_times(lama_slam* h, double ms[4], uint64_t launches[5]);
int lama_slam_map_bounds(lama_slam* h, int kind, uint32_t mn[2], uint32_t mx[2], int* patches);
int lama_slam_export_occupancy(lama_slam* h, uint32_t x0, uint32_t y0, int w, int hgt, uint16_t* occupied, uint16_t* visited, uint8_t* known);
/* occupancy == 1 only: dense window of ProbabilisticOccupancyMap cells (float log-odds, prob_tag) + Container known bit */
int lama_slam_occupancy_query(lama_slam* h, const uint32_t* cells_xy, int n, double* prob, uint8_t* flags);   /* as lama_pf_occupancy_query */
int lama_slam_distance(lama_slam* h, const double* pts_xyz, int n, double* dist, double* grad);                  /* as lama_pf_distance */
int lama_slam_write_map(lama_slam* h, int kind, const char* path);                                     /* as lama_pf_write_map */
int lama_slam_export_image(lama_slam* h, int kind, uint8_t* pixels, size_t cap, int dims[2]);           /* as lama_pf_export_image */
int lama_slam_export_logodds(lama_slam* h, uint32_t x0, uint32_t y0, int w, int hgt, float* logodds, uint8_t* known);
int lama_slam_export_distance(lama_slam* h, uint32_t x0, uint32_t y0, int w, int hgt, uint16_t* sqdist, uint8_t* valid, uint8_t* known,
                              int16_t* ox, int16_t* oy, uint8_t* queued);

/* ------------------------------------------------------------------------------------------------
 * Loc2D (match path) -- include/lama/loc2d.h:59-130, src/loc2d.cpp:46-192
 * The caller fills the public distance_map (loc2d.h:104) through the lama_dm_* grid interface.
 * ------------------------------------------------------------------------------------------------ */
typedef struct lama_loc lama_loc;
typedef struct lama_dm lama_dm; /* a device-resident DynamicDistanceMap */
typedef struct lama_loc_options { /* Loc2D::Options, loc2d.cpp:46-58 */
    double trans_thresh, rot_thresh, l2_max, resolution;
    uint32_t patch_size, max_iter;
    int32_t strategy;
    uint32_t gloc_particles, gloc_iters; /* globalLocalization: candidates per attempt, attempts (loc2d.cpp:53-54) */
    double gloc_thresh;                  /* RMSE that ends global localisation (loc2d.cpp:55) */
    double cov_blend;                    /* blend of the sampling covariance into the solver covariance (loc2d.cpp:57,199-247) */
    double center_xy[2]; /* where to centre the device map window */
    lama_device_options dev;
} lama_loc_options;
int lama_loc_options_default(lama_loc_options* o);
int lama_loc_create(const lama_loc_options* o, lama_loc** out);              /* Loc2D::Init, loc2d.cpp:61-108 */
int lama_loc_destroy(lama_loc* h);
int lama_loc_distance_map(lama_loc* h, lama_dm** dm);                        /* borrowed: Loc2D::distance_map */
int lama_loc_set_pose(lama_loc* h, const double xyr[3]);                     /* Loc2D::setPose, loc2d.h:117-118 */
int lama_loc_update(lama_loc* h, const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4],
                    const double odom_xyr[3], double timestamp, int force_update, int* did_update); /* loc2d.cpp:126-192 */
int lama_loc_get_pose(lama_loc* h, double xyr[3]);
int lama_loc_get_state(lama_loc* h, double state[4]);
int lama_loc_get_covar(lama_loc* h, double cov[9]);                          /* Loc2D::getCovar */
int lama_loc_get_rmse(lama_loc* h, double* rmse);                            /* Loc2D::getRMSE */
int lama_loc_get_solve_stats(lama_loc* h, uint32_t stats[2]);                /* {iterations, residual evaluations} */
/* public `occupancy_map` (SimpleOccupancyMap, loc2d.h:103): setFree (state -1) / setUnknown (0) / setOccupied (1) on n cells */
int lama_loc_occupancy_set(lama_loc* h, const uint32_t* cells_xy, int n, int state);
int lama_loc_occupancy_read(lama_loc* h, const char* path);                   /* occupancy_map->read(file): SimpleOccupancyMap, Map::read map.cpp:531-575 */
int lama_loc_set_seed(lama_loc* h, uint32_t seed);                            /* random::setSeed for the sampling below */
int lama_loc_trigger_global_localization(lama_loc* h);                        /* Loc2D::triggerGlobalLocalization, loc2d.cpp:194-197 */
int lama_loc_global_localization_active(lama_loc* h, int* active);

/* ------------------------------------------------------------------------------------------------
 * SDM grid interface on a device-resident DynamicDistanceMap
 * include/lama/sdm/distance_map.h:66-70, dynamic_distance_map.h:55-66
 * ------------------------------------------------------------------------------------------------ */
int lama_dm_create(double resolution, uint32_t patch_size, double l2_max, const double center_xy[2], const lama_device_options* dev, lama_dm** out);
int lama_dm_destroy(lama_dm* dm);
int lama_dm_max_sqdist(lama_dm* dm, uint32_t* max_sqdist);
/* addObstacle / removeObstacle on n cells in list order, dynamic_distance_map.cpp:212-242; nothing propagates until update */
int lama_dm_add_obstacles(lama_dm* dm, const uint32_t* cells_xy, int n);
int lama_dm_remove_obstacles(lama_dm* dm, const uint32_t* cells_xy, int n);
/* DynamicDistanceMap::update(), dynamic_distance_map.cpp:160-197; *processed = its return value */
int lama_dm_update(lama_dm* dm, uint32_t* processed);
/* DistanceMap::distance(Vector3d, Vector3d* grad) for n points; grad (n x 3) may be NULL */
int lama_dm_distance(lama_dm* dm, const double* pts_xyz, int n, double* dist, double* grad);
int lama_dm_bounds(lama_dm* dm, uint32_t mn[2], uint32_t mx[2], int* patches);
int lama_dm_write(lama_dm* dm, const char* path);   /* Map::write, map.cpp:490-529 */
int lama_dm_read(lama_dm* dm, const char* path);    /* Map::read, map.cpp:531-575, into an empty map of the same resolution and l2_max */
int lama_dm_export_image(lama_dm* dm, uint8_t* pixels, size_t cap, int dims[2]);   /* sdm::export_to_png(DistanceMap), export.cpp:75-96 */
int lama_dm_export(lama_dm* dm, uint32_t x0, uint32_t y0, int w, int hgt, uint16_t* sqdist, uint8_t* valid, uint8_t* known, int16_t* ox,
                   int16_t* oy, uint8_t* queued);
/* upload distance_t fields for a patch-aligned window (x0, y0, w, hgt multiples of 32); cells with known == 0 are left absent */
int lama_dm_import(lama_dm* dm, uint32_t x0, uint32_t y0, int w, int hgt, const uint16_t* sqdist, const uint8_t* valid, const uint8_t* known,
                   const int16_t* ox, const int16_t* oy, const uint8_t* queued);

/* Solver plug point: the weighted normal equations of MatchSurface2D at `count` SE2 states on one distance
 * map, so that a host nlls::Strategy (nlls/strategy.h:43-82) can drive the device evaluation.
 * out = count x 12: {A00,A01,A02,A11,A12,A22, g0,g1,g2, chi2, sum d^2, sum -d^2/meas_sigma} */
int lama_dm_match_normal_equations(lama_dm* dm, const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4],
                                   const double* states, int count, int robust_kind, double robust_param, double meas_sigma, double* out);
/* Solve(options, MatchSurface2D, cov) for `count` start states (nlls/solver.h:84); states updated in place;
 * stats = count x 2 {iterations, evaluations}; sums = count x 12 at the final states (may be NULL) */
int lama_dm_match_solve(lama_dm* dm, const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4], double* states,
                        int count, int strategy, int robust_kind, double robust_param, uint32_t max_iter, uint32_t* stats, double* sums);

/* MatchSurface2D::error() (src/match_surface_2d.cpp:92-116): sqrt(sum d^2 / N), d = nearest-cell distance, of the cloud at `count` states */
int lama_dm_match_error(lama_dm* dm, const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4], const double* states, int count,
                        double* rmse);

/* ---- lama::SimplePGO::optimize (include/lama/simple_pgo.h:43-57, src/simple_pgo.cpp:48-105): pose-graph optimisation on the device ----
 * nodes_xyr = node_list (n_nodes x {x, y, rotation}), edges = edge_list (from, to pairs + measured relative poses), fixed = fixed_list.
 * The factor graph is the reference's: a prior on node 0 (sigmas 1) or on every fixed node (sigmas 0.1), BetweenFactor<SE2> on consecutive
 * nodes (measured = node[i]^-1 node[i+1]) and on the edges, sigmas (0.5, 0.5, 0.1), miniSAM Levenberg-Marquardt with diagonal damping
 * (vendor/minisam/minisam/nonlinear/LevenbergMarquardtOptimizer.cpp:56-332).  *status = miniSAM's NonlinearOptimizationStatus (0 SUCCESS,
 * 1 MAX_ITERATION, 2 ERROR_INCREASE, 3 RANK_DEFICIENCY, 4 INVALID); like the reference, nodes_xyr is only updated on SUCCESS.
 * report = {LM iterations, lambda tries, CG iterations, initial error, final error, device ms}. */
int lama_pgo_optimize(int device, double* nodes_xyr, int n_nodes, const int* edges_from_to, const double* edges_xyr, int n_edges, const int* fixed_nodes,
                      const double* fixed_xyr, int n_fixed, int* status, double report[6]);

/* ---- GraphSlam2D's loop-closure front end on a device map (src/graph_slam2d.cpp:283-392) ----
 * findLoopClosureCandidates (:283-313): ids of the key poses (x, y pairs) within `radius` of the query among the first
 * n_keys - ignore_n_chain_poses, nearest first, at most max_candidates (Options::loop_max_candidates, graph_slam2d.h:75). */
int lama_loop_closure_candidates(const double* key_xy, int n_keys, int ignore_n_chain_poses, const double query_xy[2], double radius, int max_candidates, int* ids,
                                 int* count);
/* correlateCandidateScan (:315-355): the candidate key pose's cloud against the distance map -- one Gauss-Newton iteration (Huber 0.15) from
 * the candidate's pose and one from the reference position, the better start refined to convergence; between = matched pose - ref pose
 * (Pose2D::operator-, pose2d.cpp:81-84), *rmse = MatchSurface2D::error() there.  ref / cand poses are the corrected key poses (:319-320). */
int lama_slam_correlate_candidate_scan(lama_slam* h, const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4],
                                       const double ref_xyr[3], const double cand_xyr[3], double between_xyr[3], double* rmse);
int lama_dm_correlate_candidate_scan(lama_dm* dm, const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4],
                                     const double ref_xyr[3], const double cand_xyr[3], double between_xyr[3], double* rmse);
/* coarseSearchAndCorrelateCandidateScan (:357-392): first against a coarse distance map (0.25 m cells, 2.5 m reach) built from the reference
 * key pose's cloud alone, then against the map itself. */
int lama_slam_coarse_correlate_candidate_scan(lama_slam* h, const double* ref_pts_xyz, int ref_n, const double ref_origin[3], const double ref_quat_xyzw[4],
                                              const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4],
                                              const double ref_xyr[3], const double cand_xyr[3], double between_xyr[3], double* rmse);
int lama_dm_coarse_correlate_candidate_scan(lama_dm* dm, const double* ref_pts_xyz, int ref_n, const double ref_origin[3], const double ref_quat_xyzw[4],
                                            const double* pts_xyz, int n, const double sensor_origin[3], const double sensor_quat_xyzw[4],
                                            const double ref_xyr[3], const double cand_xyr[3], double between_xyr[3], double* rmse);

#ifdef __cplusplus
}
#endif
#endif /* LAMA_B200_H */
