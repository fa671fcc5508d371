// ocm.cu -- keyframe RGB-D -> point cloud -> occupancy (OctoMap log-odds) on sm_100a.
// Reference: MapDrawer::GeneratePointCloud / InsertScan / ctor params (perfect/src/MapDrawer.cc:641-675, 946-1025,
// 51-56), UpdateOctomap's sensor-origin quirk (:619,631-632), PointCloudMapping::insertKeyFrame (class surface,
// include/pointcloudmapping.h:50-56).  Third-party semantics (PCL VoxelGrid, pcl::transformPointCloud, octomap
// keys / computeRayKeys / updateNode) follow oracle/occ_ref.cpp, which states them.
//
// HBM layout:
//   leaf table   open-addressing hash of the 1-cm VoxelGrid cells of ONE keyframe (key = 3 x 21-bit cell index), one
//                32-byte record per cell: key, count, first pixel, bucket offset, cursor.  Pixels of a cell are bucketed, sorted ascending and
//                summed sequentially in float so the centroid equals the oracle's pixel-order sum bit for bit.
//   points       world-frame centroids of the last keyframe (xyz f32, rgb, ground label), unordered
//   map          persistent open-addressing hash: key (3 x u16 OcTreeKey packed in u64) -> float log-odds, colour, the
//                clamp-add summary (a, lo, hi) used by the multi-GPU merge, and the batch masks: bit j of the hit / miss
//                word = keyframe j of the running batch observed the voxel occupied / free.  The scan of ALL keyframes
//                of a batch only ORs bits (order-free, one launch); one more launch replays each touched voxel's bits in
//                keyframe order, which reproduces the sequential clamped updates exactly.
#include <algorithm>
#include <new>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace b200 {

constexpr unsigned long long EMPTY_KEY = ~0ull;

__device__ __forceinline__ unsigned long long hash64(unsigned long long k) {   // splitmix64 finaliser
  k ^= k >> 30; k *= 0xbf58476d1ce4e5b9ull;
  k ^= k >> 27; k *= 0x94d049bb133111ebull;
  k ^= k >> 31;
  return k;
}
// Home slot of a key in a power-of-two table.  Tables below 2^32 slots (every one in practice) take a 32-bit mix of the
// two key halves (murmur3 block + finaliser): about half the instructions of the 64-bit multiplies, and the probe is on
// the inner loop of the free-space rays (k_ocm_scan_keys: the 64-bit hash was 19 % of its instructions).
__device__ __forceinline__ long long hash_slot(unsigned long long k, long long mask) {
  if ((unsigned long long)mask >> 32) return (long long)(hash64(k) & (unsigned long long)mask);
  unsigned h = (unsigned)k * 0xcc9e2d51u;
  h = __funnelshift_l(h, h, 15) * 0x1b873593u;
  h ^= (unsigned)(k >> 32) * 0x85ebca6bu;
  h ^= h >> 16; h *= 0x85ebca6bu;
  h ^= h >> 13; h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return (long long)(h & (unsigned)mask);
}

// find-or-insert; returns slot or -1 when the table is full; *fresh = this call created the entry
__device__ __forceinline__ long long table_insert(unsigned long long* keys, long long cap_mask, unsigned long long key,
                                                  bool* fresh) {
  long long s = hash_slot(key, cap_mask);
  *fresh = false;
  for (long long probe = 0; probe <= cap_mask; ++probe) {
    const unsigned long long cur = keys[s];
    if (cur == key) return s;
    if (cur == EMPTY_KEY) {
      const unsigned long long old = atomicCAS(&keys[s], EMPTY_KEY, key);
      if (old == EMPTY_KEY) { *fresh = true; return s; }
      if (old == key) return s;
    }
    s = (s + 1) & cap_mask;
  }
  return -1;
}

__device__ __forceinline__ long long table_find(const unsigned long long* keys, long long cap_mask, unsigned long long key) {
  long long s = hash_slot(key, cap_mask);
  for (long long probe = 0; probe <= cap_mask; ++probe) {
    const unsigned long long cur = keys[s];
    if (cur == key) return s;
    if (cur == EMPTY_KEY) return -1;
    s = (s + 1) & cap_mask;
  }
  return -1;
}

struct OcmConst {
  float fx, fy, cx, cy;
  float depth_min, depth_max, y_max, leaf, inv_leaf;
  double res, res_factor;
  float hit_log, miss_log, cmin, cmax;
  double Rt[9], ti[3];     // Twc = inverse of Tcw, double (pcl::transformPointCloud<.., double>)
  float origin[3];         // translation of Tcw (UpdateOctomap quirk)
  int rows, cols;
};

// One 32-byte record (= one L2 sector) per VoxelGrid cell: an insert, the bucket bookkeeping and the reset all touch
// the same sector instead of five arrays.
struct __align__(32) LeafRec {
  unsigned long long key;
  int count;
  int first;     // smallest pixel index
  int offset;
  int cursor;
  int pad[2];
};
struct LeafTable {
  LeafRec* rec;
  long long mask;
};

// find-or-insert on the record table; returns slot or -1 when the table is full
__device__ __forceinline__ long long leaf_insert(const LeafTable& lt, unsigned long long key, bool* fresh) {
  long long s = hash_slot(key, lt.mask);
  *fresh = false;
  for (long long probe = 0; probe <= lt.mask; ++probe) {
    const unsigned long long cur = lt.rec[s].key;
    if (cur == key) return s;
    if (cur == EMPTY_KEY) {
      const unsigned long long old = atomicCAS(&lt.rec[s].key, EMPTY_KEY, key);
      if (old == EMPTY_KEY) { *fresh = true; return s; }
      if (old == key) return s;
    }
    s = (s + 1) & lt.mask;
  }
  return -1;
}

struct MapView {
  unsigned long long* keys;
  float* val;
  float *a, *lo, *hi;   // clamp-add summary since the last reset: f(x) = min(max(x + a, lo), hi)
  unsigned* rgb;
  int* nleaves;
  long long mask;
  // batch accumulation: which keyframes of the running batch hit (high word) / missed (low word) the voxel, the colour
  // of its latest hit (keyframe << 24 | rgb); k_ocm_apply finds the touched voxels by sweeping the masks
  unsigned long long* bm;
  unsigned* bm_rgb;
  // merge epochs (ocm_merge_nccl): value of the voxel at the last merge, and the list of voxels whose summary left the
  // identity since then (what this rank has to send)
  float* base;
  int* elist;
  int* nelist;
  unsigned long long* nupd;   // cumulative number of (voxel, keyframe) log-odds updates (bench.py: U of B_map)
};

// Record "keyframe j of the batch observed voxel `key` as occupied / free".  Order-free (atomicOr), so every keyframe
// of a batch can be scanned in one launch; k_ocm_apply replays the bits of each voxel in keyframe order.
__device__ __forceinline__ bool map_touch(const MapView& m, unsigned long long key, int j, bool occupied, unsigned rgb) {
  long long s = hash_slot(key, m.mask);
  bool found = false;
  for (long long probe = 0; probe <= m.mask; ++probe) {
    const unsigned long long cur = m.keys[s];
    if (cur == key) { found = true; break; }
    if (cur == EMPTY_KEY) {
      const unsigned long long old = atomicCAS(&m.keys[s], EMPTY_KEY, key);
      if (old == EMPTY_KEY) { atomicAdd(m.nleaves, 1); found = true; break; }
      if (old == key) { found = true; break; }
    }
    s = (s + 1) & m.mask;
  }
  if (!found) return false;
  // Fire and forget: no atomic here returns a value.  (The free-space rays used to wait on the atomicOr's old value to
  // build a list of touched voxels: 35 % of k_ocm_scan_keys' stall samples.  Dirty-block flags instead of the list were
  // worse: the hash scatters a batch over every block, and millions of flag stores / flag reads land on 64 cache lines.
  // k_ocm_apply simply sweeps the 8-byte masks of the whole table: 67 MB per batch at the default capacity.)
  const unsigned long long bit = occupied ? (1ull << (32 + j)) : (1ull << j);
  atomicOr(&m.bm[s], bit);
  if (occupied) atomicMax(&m.bm_rgb[s], ((unsigned)j << 24) | (rgb & 0xffffffu));
  return true;
}

// Everything one keyframe needs between back-projection and the map update.  A batch of keyframes runs every stage
// as ONE launch (blockIdx.y = job = bit index of the keyframe in the batch masks of the map).
struct KfScratch {
  LeafTable leaf;
  int *pix_slot, *pix_rank, *bucket, *voxlist;   // pix_rank: position of the pixel inside its cell's bucket
  int* counters;   // [0] pixels, [1] voxels (= points)
  float* pts;
  uint8_t *pts_rgb, *pts_label;
};
struct KfJob {
  OcmConst c;
  const float* depth;
  const uint8_t* rgb;
  const uint8_t* label;
  KfScratch s;
};

__device__ __forceinline__ bool backproject(const OcmConst& c, const float* __restrict__ depth, int pix, float& x,
                                            float& y, float& z) {
  const int m = pix / c.cols, n = pix - m * c.cols;
  const float d = depth[pix];
  if (d < c.depth_min || d > c.depth_max) return false;       // MapDrawer.cc:655 (NaN fails neither test: kept like the reference)
  z = d;
  x = ((float)n - c.cx) * z / c.fx;                           // :658-659 (float, true division; --fmad=false)
  y = ((float)m - c.cy) * z / c.fy;
  if (y < -c.y_max || y > c.y_max) return false;             // :660
  return true;
}

__device__ __forceinline__ unsigned long long leaf_key(const OcmConst& c, float x, float y, float z) {
  const long long ix = (long long)floorf(x * c.inv_leaf) + (1 << 20);
  const long long iy = (long long)floorf(y * c.inv_leaf) + (1 << 20);
  const long long iz = (long long)floorf(z * c.inv_leaf) + (1 << 20);
  return (unsigned long long)(ix & 0x1fffff) | ((unsigned long long)(iy & 0x1fffff) << 21) |
         ((unsigned long long)(iz & 0x1fffff) << 42);
}

// Pixel of this thread for the warp-tiled kernels: a warp covers an 8 x 4 pixel tile (lane = ty * 8 + tx, so lanes are
// in row-major order inside the tile).  Neighbouring pixels mostly fall into the same 1-cm cell, so the cell-level
// atomics are issued once per (warp, cell) by the lowest lane of each __match_any group.
__device__ __forceinline__ int tile_pixel(int rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int tiles_x = (cols + 7) >> 3;
  const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int y = ty * 4 + (lane >> 3), x = tx * 8 + (lane & 7);
  return (y < rows && x < cols) ? y * cols + x : -1;
}
__host__ __device__ inline int tile_blocks(int rows, int cols, int warps_per_block) {
  const int tiles = ((cols + 7) >> 3) * ((rows + 3) >> 2);
  return (tiles + warps_per_block - 1) / warps_per_block;
}

// K11a: gate + VoxelGrid cell of every pixel
__global__ void k_ocm_bin(const KfJob* __restrict__ jobs, int* __restrict__ err) {
  const KfJob& J = jobs[blockIdx.y];
  const OcmConst& c = J.c;
  const float* __restrict__ depth = J.depth;
  const LeafTable lt = J.s.leaf;
  int* __restrict__ pix_slot = J.s.pix_slot;
  int* __restrict__ counters = J.s.counters;
  int* __restrict__ voxlist = J.s.voxlist;
  const int lane = threadIdx.x & 31;
  const int pix = tile_pixel(c.rows, c.cols);
  float x, y, z;
  const bool ok = pix >= 0 && backproject(c, depth, pix, x, y, z);
  const unsigned long long key = ok ? leaf_key(c, x, y, z) : (EMPTY_KEY - 1ull - (unsigned long long)lane);
  const unsigned grp = __match_any_sync(0xffffffffu, key);
  const int leader = __ffs(grp) - 1;
  int slot = -1, rank = 0;
  if (ok && lane == leader) {   // lowest lane of the group = smallest pixel index of the group
    bool fresh;
    const long long s = leaf_insert(lt, key, &fresh);
    if (s < 0) { atomicExch(err, 1); }
    else {
      slot = (int)s;
      if (fresh) voxlist[atomicAdd(&counters[1], 1)] = slot;   // list of occupied cells: no table sweep later
      rank = atomicAdd(&lt.rec[s].count, __popc(grp));          // the group's first position in the cell's bucket
      atomicMin(&lt.rec[s].first, pix);
    }
  }
  slot = __shfl_sync(0xffffffffu, slot, leader);
  rank = __shfl_sync(0xffffffffu, rank, leader) + __popc(grp & ((1u << lane) - 1u));
  if (pix >= 0) {
    pix_slot[pix] = ok ? slot : -1;
    J.s.pix_rank[pix] = rank;
  }
}

// K11b: bucket ranges for the occupied cells (arbitrary order: the cloud is a set)
__global__ void k_ocm_ranges(const KfJob* __restrict__ jobs) {
  const KfJob& J = jobs[blockIdx.y];
  const LeafTable lt = J.s.leaf;
  int* __restrict__ counters = J.s.counters;   // [0]=pixels, [1]=voxels
  const int* __restrict__ voxlist = J.s.voxlist;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= counters[1]) return;
  const int s = voxlist[v];
  lt.rec[s].offset = atomicAdd(&counters[0], lt.rec[s].count);
}

// K11c: scatter the pixel indices into their cell's bucket
__global__ void k_ocm_scatter(const KfJob* __restrict__ jobs) {
  const KfJob& J = jobs[blockIdx.y];
  const int pix = tile_pixel(J.c.rows, J.c.cols);
  if (pix < 0) return;
  const int s = J.s.pix_slot[pix];
  if (s >= 0) J.s.bucket[J.s.leaf.rec[s].offset + J.s.pix_rank[pix]] = pix;   // positions were handed out by k_ocm_bin
}

// K11d: per cell: restore pixel order, sequential float centroid (PCL VoxelGrid), transform to the world frame in
// double (pcl::transformPointCloud), emit the point; also resets the cell's table entry for the next keyframe.
__global__ void k_ocm_centroids(const KfJob* __restrict__ jobs) {
  const KfJob& J = jobs[blockIdx.y];
  const OcmConst& c = J.c;
  const float* __restrict__ depth = J.depth;
  const uint8_t* __restrict__ rgb = J.rgb;
  const uint8_t* __restrict__ label = J.label;
  const LeafTable lt = J.s.leaf;
  const int* __restrict__ counters = J.s.counters;
  const int* __restrict__ voxlist = J.s.voxlist;
  int* __restrict__ bucket = J.s.bucket;
  float* __restrict__ pts = J.s.pts;
  uint8_t* __restrict__ pts_rgb = J.s.pts_rgb;
  uint8_t* __restrict__ pts_label = J.s.pts_label;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= counters[1]) return;
  const int s = voxlist[v];
  const int n = lt.rec[s].count;
  int* b = bucket + lt.rec[s].offset;
  for (int i = 1; i < n; ++i) {   // insertion sort: ascending pixel index = row-major order
    const int val = b[i];
    int j = i - 1;
    while (j >= 0 && b[j] > val) { b[j + 1] = b[j]; --j; }
    b[j + 1] = val;
  }
  float sx = 0.f, sy = 0.f, sz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
  for (int i = 0; i < n; ++i) {
    float x, y, z;
    backproject(c, depth, b[i], x, y, z);
    sx += x; sy += y; sz += z;
    const uint8_t* col = rgb + (size_t)b[i] * 3;   // b,g,r (cv::Mat BGR)
    sr += (float)col[2]; sg += (float)col[1]; sb += (float)col[0];
  }
  const float fn = (float)n;
  const double px = (double)(sx / fn), py = (double)(sy / fn), pz = (double)(sz / fn);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    pts[(size_t)v * 3 + i] = (float)(c.Rt[i * 3 + 0] * px + c.Rt[i * 3 + 1] * py + c.Rt[i * 3 + 2] * pz + c.ti[i]);
  pts_rgb[(size_t)v * 3 + 0] = (uint8_t)(sr / fn);
  pts_rgb[(size_t)v * 3 + 1] = (uint8_t)(sg / fn);
  pts_rgb[(size_t)v * 3 + 2] = (uint8_t)(sb / fn);
  pts_label[v] = label ? label[lt.rec[s].first] : 0;
  LeafRec fresh_rec;
  fresh_rec.key = EMPTY_KEY; fresh_rec.count = 0; fresh_rec.first = 0x7fffffff; fresh_rec.offset = 0; fresh_rec.cursor = 0;
  fresh_rec.pad[0] = fresh_rec.pad[1] = 0;
  lt.rec[s] = fresh_rec;
}

// leaf <= 0: no VoxelGrid; every gated pixel is a point (used for the T-variant style clouds)
__global__ void k_ocm_points_nofilter(const KfJob* __restrict__ jobs) {
  const KfJob& J = jobs[blockIdx.y];
  const OcmConst& c = J.c;
  const float* __restrict__ depth = J.depth;
  const uint8_t* __restrict__ rgb = J.rgb;
  const uint8_t* __restrict__ label = J.label;
  int* __restrict__ counters = J.s.counters;
  float* __restrict__ pts = J.s.pts;
  uint8_t* __restrict__ pts_rgb = J.s.pts_rgb;
  uint8_t* __restrict__ pts_label = J.s.pts_label;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= c.rows * c.cols) return;
  float x, y, z;
  if (!backproject(c, depth, pix, x, y, z)) return;
  const int v = atomicAdd(&counters[1], 1);
  const double px = x, py = y, pz = z;
#pragma unroll
  for (int i = 0; i < 3; ++i)
    pts[(size_t)v * 3 + i] = (float)(c.Rt[i * 3 + 0] * px + c.Rt[i * 3 + 1] * py + c.Rt[i * 3 + 2] * pz + c.ti[i]);
  const uint8_t* col = rgb + (size_t)pix * 3;
  pts_rgb[(size_t)v * 3 + 0] = col[2]; pts_rgb[(size_t)v * 3 + 1] = col[1]; pts_rgb[(size_t)v * 3 + 2] = col[0];
  pts_label[v] = label ? label[pix] : 0;
}

__device__ __forceinline__ bool coord_to_key(const OcmConst& c, float coord, int& k) {   // coordToKeyChecked
  const int s = (int)floor(c.res_factor * (double)coord) + 32768;
  k = s;
  return s >= 0 && s < 65536;
}
__device__ __forceinline__ unsigned long long pack_key(int kx, int ky, int kz) {
  return (unsigned long long)kx | ((unsigned long long)ky << 16) | ((unsigned long long)kz << 32);
}

// K12/K13: per point -> occupied endpoint key, or (ground) the free cells along the ray (computeRayKeys); the
// observations go straight into the map's batch masks (bit = keyframe).
// Free-space rays of neighbouring ground points run through the same cells for most of their length (they share the
// origin), and a miss is idempotent inside one keyframe (atomicOr of the same bit): every CTA keeps a small direct-mapped
// cache of the free cells it has already reported and skips the repeats -- the global hash probes + atomics were the whole
// kernel (long-scoreboard bound), most of them redundant.
constexpr int SCAN_CACHE = 2048;
__device__ __forceinline__ bool scan_touch_free(const MapView& m, unsigned long long* seen, unsigned long long key, int j) {
  const unsigned h = ((unsigned)key * 73856093u ^ (unsigned)(key >> 16) * 19349663u ^ (unsigned)(key >> 32) * 83492791u) &
                     (SCAN_CACHE - 1);
  if (seen[h] == key) return true;
  seen[h] = key;   // benign race: at worst a repeat reaches the map
  return map_touch(m, key, j, false, 0u);
}

__global__ void __launch_bounds__(256) k_ocm_scan_keys(const KfJob* __restrict__ jobs, MapView m, int* __restrict__ err) {
  __shared__ unsigned long long s_seen[SCAN_CACHE];
  for (int i = threadIdx.x; i < SCAN_CACHE; i += blockDim.x) s_seen[i] = EMPTY_KEY;
  __syncthreads();
  const int j = blockIdx.y;
  const KfJob& J = jobs[j];
  const OcmConst& c = J.c;
  const int* __restrict__ counters = J.s.counters;
  const float* __restrict__ pts = J.s.pts;
  const uint8_t* __restrict__ pts_label = J.s.pts_label;
  const uint8_t* __restrict__ pts_rgb = J.s.pts_rgb;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= counters[1]) return;
  // perfect/src/MapDrawer.cc:676-680: a cloud of fewer than 50 points skips the plane extraction and is ALL ground
  const bool all_ground = counters[1] < 50;
  const float e[3] = {pts[(size_t)v * 3], pts[(size_t)v * 3 + 1], pts[(size_t)v * 3 + 2]};
  int ke[3];
  const bool end_ok = coord_to_key(c, e[0], ke[0]) && coord_to_key(c, e[1], ke[1]) && coord_to_key(c, e[2], ke[2]);
  if (!pts_label[v] && !all_ground) {
    if (end_ok) {
      const unsigned rgb = (unsigned)pts_rgb[(size_t)v * 3] | ((unsigned)pts_rgb[(size_t)v * 3 + 1] << 8) |
                           ((unsigned)pts_rgb[(size_t)v * 3 + 2] << 16);
      if (!map_touch(m, pack_key(ke[0], ke[1], ke[2]), j, true, rgb)) atomicExch(err, 4);
    }
    return;
  }
  int ko[3];
  if (!end_ok || !coord_to_key(c, c.origin[0], ko[0]) || !coord_to_key(c, c.origin[1], ko[1]) ||
      !coord_to_key(c, c.origin[2], ko[2]))
    return;
  if (ko[0] == ke[0] && ko[1] == ke[1] && ko[2] == ke[2]) return;
  if (!scan_touch_free(m, s_seen, pack_key(ko[0], ko[1], ko[2]), j)) atomicExch(err, 4);
  float dir[3] = {e[0] - c.origin[0], e[1] - c.origin[1], e[2] - c.origin[2]};
  // octomath::Vector3::norm() of octomap 1.9.x: norm_sq() = x*x + y*y + z*z in float, widened for the sqrt only
  const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(dir[0], dir[0]), __fmul_rn(dir[1], dir[1])), __fmul_rn(dir[2], dir[2]));
  const float length = (float)sqrt((double)n2);
#pragma unroll
  for (int i = 0; i < 3; ++i) dir[i] = dir[i] / length;
  int step[3], cur[3] = {ko[0], ko[1], ko[2]};
  double tMax[3], tDelta[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    step[i] = (dir[i] > 0.f) ? 1 : ((dir[i] < 0.f) ? -1 : 0);
    if (step[i] != 0) {
      double border = ((double)(cur[i] - 32768) + 0.5) * c.res;
      border += (double)(float)((double)step[i] * c.res * 0.5);
      tMax[i] = (border - (double)c.origin[i]) / (double)dir[i];
      tDelta[i] = c.res / fabs((double)dir[i]);
    } else {
      tMax[i] = 1.7976931348623157e308;
      tDelta[i] = 1.7976931348623157e308;
    }
  }
  // computeRayKeys' loop (octomap OcTreeBaseImpl.hxx:  step along the axis of the smallest tMax; stop at the end key;
  // stop when min(tMax) after the step exceeds the ray length; else the cell is free).  min(tMax) after a step IS the
  // tMax of the axis the next step takes, so the distance test of cell k rides on the axis selection of step k+1 and
  // the cell is reported one iteration late (`pending`): no separate three-way double minimum per cell.
  bool pending = false;
  unsigned long long pending_key = 0ull;
  const double dlen = (double)length;
  for (int guard = 0; guard < 3 * 65536; ++guard) {
    const bool a01 = tMax[0] < tMax[1];
    const double t01 = a01 ? tMax[0] : tMax[1];
    const bool a2 = t01 < tMax[2];            // (tMax[0] < tMax[2]) or (tMax[1] < tMax[2]), whichever axis leads
    const double tmin = a2 ? t01 : tMax[2];
    if (pending) {
      if (tmin > dlen) break;
      if (!scan_touch_free(m, s_seen, pending_key, j)) { atomicExch(err, 4); break; }
    }
    const bool d0 = a2 && a01, d1 = a2 && !a01, d2 = !a2;
    // (no dynamic register indexing, no branches: predicated updates)
    cur[0] = d0 ? ((cur[0] + step[0]) & 0xffff) : cur[0];
    cur[1] = d1 ? ((cur[1] + step[1]) & 0xffff) : cur[1];
    cur[2] = d2 ? ((cur[2] + step[2]) & 0xffff) : cur[2];
    tMax[0] = d0 ? tMax[0] + tDelta[0] : tMax[0];
    tMax[1] = d1 ? tMax[1] + tDelta[1] : tMax[1];
    tMax[2] = d2 ? tMax[2] + tDelta[2] : tMax[2];
    if (cur[0] == ke[0] && cur[1] == ke[1] && cur[2] == ke[2]) break;
    pending = true;
    pending_key = pack_key(cur[0], cur[1], cur[2]);
  }
}

// K14: replay, voxel by voxel, the observations of the batch in keyframe order (MapDrawer.cc:1007-1022 per keyframe:
// free \ occupied get a miss, occupied get a hit; updateNodeLogOdds clamps after every add, so the order matters and
// is kept).  Grid-stride sweep over the batch masks of the whole table (16-byte loads); the masks of the voxels found
// are cleared for the next batch.
__device__ __forceinline__ int apply_slot(const MapView& m, long long s, unsigned long long w, float hit_log, float miss_log,
                                          float cmin, float cmax) {
  m.bm[s] = 0ull;
  const unsigned hit = (unsigned)(w >> 32), miss = (unsigned)w & ~hit;
  float v = m.val[s], a = m.a[s], lo = m.lo[s], hi = m.hi[s];
  if (a == 0.f && lo == -INFINITY) m.elist[atomicAdd(m.nelist, 1)] = (int)s;   // first update of this merge epoch
  unsigned all = hit | miss;
  const int nup = __popc(all);
  while (all) {
    const int j = __ffs(all) - 1;
    all &= all - 1u;
    const float d = ((hit >> j) & 1u) ? hit_log : miss_log;
    v = fminf(fmaxf(v + d, cmin), cmax);
    a = a + d;
    lo = fminf(fmaxf(lo + d, cmin), cmax);
    hi = fminf(fmaxf(hi + d, cmin), cmax);
  }
  m.val[s] = v; m.a[s] = a; m.lo[s] = lo; m.hi[s] = hi;
  if (hit) { m.rgb[s] = (m.bm_rgb[s] & 0xffffffu) | 0x01000000u; m.bm_rgb[s] = 0u; }   // top byte 1: coloured this epoch
  return nup;
}

__global__ void k_ocm_apply(float hit_log, float miss_log, float cmin, float cmax, MapView m) {
  const long long C = m.mask + 1;   // power of two >= 2
  const ulonglong2* __restrict__ bm2 = reinterpret_cast<const ulonglong2*>(m.bm);
  int nup = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < C / 2; i += (long long)gridDim.x * blockDim.x) {
    const ulonglong2 w = bm2[i];
    if (w.x) nup += apply_slot(m, 2 * i, w.x, hit_log, miss_log, cmin, cmax);
    if (w.y) nup += apply_slot(m, 2 * i + 1, w.y, hit_log, miss_log, cmin, cmax);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nup += __shfl_xor_sync(0xffffffffu, nup, o);
  if ((threadIdx.x & 31) == 0 && nup) atomicAdd(m.nupd, (unsigned long long)nup);
}

__global__ void k_ocm_fill_leaf(LeafRec* p, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  LeafRec r;
  r.key = EMPTY_KEY; r.count = 0; r.first = 0x7fffffff; r.offset = 0; r.cursor = 0; r.pad[0] = r.pad[1] = 0;
  p[i] = r;
}
__global__ void k_ocm_fill_u64(unsigned long long* p, unsigned long long v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_ocm_fill_i32(int* p, int v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void k_ocm_fill_f32(float* p, float v, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// export: compact the map (or its summaries) into dense arrays
__global__ void k_ocm_export(MapView m, long long* __restrict__ counter, long long cap, unsigned short* __restrict__ keys3,
                             unsigned long long* __restrict__ keys64, float* __restrict__ val, unsigned char* __restrict__ rgb,
                             float* __restrict__ a, float* __restrict__ lo, float* __restrict__ hi, int only_touched) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > m.mask) return;
  const unsigned long long k = m.keys[i];
  if (k == EMPTY_KEY) return;
  if (only_touched && m.a[i] == 0.f && m.lo[i] == -INFINITY) return;
  const long long o = (long long)atomicAdd((unsigned long long*)counter, 1ull);
  if (o >= cap) return;
  if (keys3) { keys3[o * 3] = (unsigned short)(k & 0xffff); keys3[o * 3 + 1] = (unsigned short)((k >> 16) & 0xffff); keys3[o * 3 + 2] = (unsigned short)((k >> 32) & 0xffff); }
  if (keys64) keys64[o] = k;
  if (val) val[o] = m.val[i];
  if (rgb) { const unsigned c = m.rgb[i]; rgb[o * 3] = c & 0xff; rgb[o * 3 + 1] = (c >> 8) & 0xff; rgb[o * 3 + 2] = (c >> 16) & 0xff; }
  if (a) { a[o] = m.a[i]; lo[o] = m.lo[i]; hi[o] = m.hi[i]; }
}

// merge: apply a later shard's per-voxel clamp-add summaries onto this map
__global__ void k_ocm_apply_summaries(OcmConst c, MapView m, const unsigned long long* __restrict__ keys,
                                      const float* __restrict__ a, const float* __restrict__ lo,
                                      const float* __restrict__ hi, long long n, int* __restrict__ err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  long long s = hash_slot(key, m.mask);
  bool found = false;
  for (long long probe = 0; probe <= m.mask; ++probe) {
    const unsigned long long cur = m.keys[s];
    if (cur == key) { found = true; break; }
    if (cur == EMPTY_KEY) {
      const unsigned long long old = atomicCAS(&m.keys[s], EMPTY_KEY, key);
      if (old == EMPTY_KEY) { atomicAdd(m.nleaves, 1); found = true; break; }
      if (old == key) { found = true; break; }
    }
    s = (s + 1) & m.mask;
  }
  if (!found) { atomicExch(err, 4); return; }
  // value: v <- g(v); summary: f <- g o f  with g = (a, lo, hi)
  const float ga = a[i], gl = lo[i], gh = hi[i];
  if (m.a[s] == 0.f && m.lo[s] == -INFINITY) m.elist[atomicAdd(m.nelist, 1)] = (int)s;
  m.val[s] = fminf(fmaxf(m.val[s] + ga, gl), gh);
  m.a[s] = m.a[s] + ga;
  m.lo[s] = fminf(fmaxf(m.lo[s] + ga, gl), gh);
  m.hi[s] = fminf(fmaxf(m.hi[s] + ga, gl), gh);
}

// ---- epoch merge (ocm_merge_nccl) -------------------------------------------------------------------------------------
// One shard = SoA records {key u64, a, lo, hi f32, rgb u32} of the voxels a rank updated since the last merge (24 B each).
struct MergeShard {
  unsigned long long* keys;
  float *a, *lo, *hi;
  unsigned* rgb;
};
__host__ __device__ inline MergeShard merge_shard_at(void* region, long long n) {
  MergeShard r;
  char* p = (char*)region;
  r.keys = (unsigned long long*)p; p += 8 * n;
  r.a = (float*)p; p += 4 * n;
  r.lo = (float*)p; p += 4 * n;
  r.hi = (float*)p; p += 4 * n;
  r.rgb = (unsigned*)p;
  return r;
}
// pack this rank's epoch summaries and roll its own voxels back to the value of the last merge: afterwards EVERY rank
// replays ALL shards (its own included) in rank order on the same base values, so all ranks end with identical maps
__global__ void k_merge_pack(MapView m, int n, MergeShard o) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = m.elist[i];
  o.keys[i] = m.keys[s];
  o.a[i] = m.a[s]; o.lo[i] = m.lo[s]; o.hi[i] = m.hi[s];
  const unsigned c = m.rgb[s];
  const bool fresh = (c >> 24) == 1u;
  o.rgb[i] = fresh ? (c & 0xffffffu) : 0xffffffffu;
  if (fresh) m.rgb[s] = c & 0xffffffu;
  m.val[s] = m.base[s];
  m.a[s] = 0.f; m.lo[s] = -INFINITY; m.hi[s] = INFINITY;
}
// apply one shard: v <- min(max(v + a, lo), hi) per voxel (keys of a shard are distinct; shards run as consecutive launches)
__global__ void k_merge_apply(MapView m, long long n, MergeShard g, int* __restrict__ slots, int* __restrict__ err) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = g.keys[i];
  long long s = hash_slot(key, m.mask);
  bool found = false;
  for (long long probe = 0; probe <= m.mask; ++probe) {
    const unsigned long long cur = m.keys[s];
    if (cur == key) { found = true; break; }
    if (cur == EMPTY_KEY) {
      const unsigned long long old = atomicCAS(&m.keys[s], EMPTY_KEY, key);
      if (old == EMPTY_KEY) { atomicAdd(m.nleaves, 1); found = true; break; }
      if (old == key) { found = true; break; }
    }
    s = (s + 1) & m.mask;
  }
  if (!found) { atomicExch(err, 4); slots[i] = -1; return; }
  m.val[s] = fminf(fmaxf(m.val[s] + g.a[i], g.lo[i]), g.hi[i]);
  const unsigned c = g.rgb[i];
  if (c != 0xffffffffu) m.rgb[s] = c;
  slots[i] = (int)s;
}
__global__ void k_merge_commit(MapView m, long long n, const int* __restrict__ slots) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = slots[i];
  if (s >= 0) m.base[s] = m.val[s];
}

__global__ void k_ocm_query(MapView m, unsigned long long key, float* out, int* found) {
  const long long s = table_find(m.keys, m.mask, key);
  *found = s >= 0;
  *out = s >= 0 ? m.val[s] : 0.f;
}

}  // namespace b200

using namespace b200;

struct ocm {
  OcmParams prm{};
  int device = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  float hit_log, miss_log, cmin, cmax;
  // map
  MapView map{};
  long long map_cap = 0;
  // per-keyframe scratch: one slot per keyframe of a batch (grown on demand, self-cleaning between uses)
  int rows = 0, cols = 0;
  static constexpr int MAX_SLOTS = 32;
  std::vector<KfScratch> slots;
  int* d_counters = nullptr;          // 4 ints per slot
  int* d_err = nullptr;
  KfJob* d_jobs = nullptr;            // device copy of the running batch, JOB_RING batches deep
  KfJob* h_jobs = nullptr;            // page-locked staging of the same
  static constexpr int JOB_RING = 8;
  cudaEvent_t job_ev[JOB_RING] = {};
  int job_head = 0;
  int last_slot = 0;
  float* d_depth = nullptr;
  uint8_t *d_rgb = nullptr, *d_label = nullptr;
  uint16_t* d_kf_d16 = nullptr;   // staging of ocm_insert_keyframes_u16
  float* d_kf_depth = nullptr;
  uint8_t* d_kf_rgb = nullptr;
  uint8_t* d_kf_label = nullptr;
  size_t kf_cap = 0;
  long long* d_export_counter = nullptr;
  // ocm_merge_nccl staging (grown on demand)
  void *mg_send = nullptr, *mg_recv = nullptr;
  int* mg_slots = nullptr;
  int* mg_counts = nullptr;        // device, one int per rank
  int* mg_counts_h = nullptr;      // page-locked mirror
  long long mg_send_cap = 0, mg_recv_cap = 0;
  int mg_world_cap = 0;

  ~ocm() {
    DeviceGuard g(device);
    auto F = [](void* p) { if (p) cudaFree(p); };
    F(map.keys); F(map.val); F(map.a); F(map.lo); F(map.hi); F(map.rgb); F(map.nleaves); F(map.bm); F(map.bm_rgb);
    F(map.base); F(map.elist); F(map.nelist); F(map.nupd);
    F(mg_send); F(mg_recv); F(mg_slots); F(mg_counts);
    if (mg_counts_h) cudaFreeHost(mg_counts_h);
    free_scratch();
    F(d_counters); F(d_err); F(d_export_counter); F(d_kf_d16); F(d_kf_depth); F(d_kf_rgb); F(d_kf_label); F(d_jobs);
    if (h_jobs) cudaFreeHost(h_jobs);
    for (cudaEvent_t e : job_ev) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
  }
  static void free_slot(KfScratch& k) {
    auto F = [](void* p) { if (p) cudaFree(p); };
    F(k.leaf.rec); F(k.pix_slot); F(k.pix_rank); F(k.bucket); F(k.voxlist);
    F(k.pts); F(k.pts_rgb); F(k.pts_label);
    k = KfScratch{};
  }
  void free_scratch() {
    auto F = [](void* p) { if (p) cudaFree(p); };
    for (KfScratch& k : slots) free_slot(k);
    slots.clear();
    F(d_depth); F(d_rgb); F(d_label);
    d_depth = nullptr; d_rgb = d_label = nullptr;
  }
  template <class T>
  int fill(T* p, T v, long long n);
  int ensure_scratch(int r, int c, int nslots = 1);
  int insert_batch(int n, const float* const* dd, const uint8_t* const* drgb, const uint8_t* const* dlabel, int r, int c,
                   const float* Tcw, float fx, float fy, float cx, float cy);
  int insert(const float* dd, const uint8_t* drgb, const uint8_t* dlabel, int r, int c, const float* Tcw, float fx,
             float fy, float cx, float cy) {
    return insert_batch(1, &dd, &drgb, &dlabel, r, c, Tcw, fx, fy, cx, cy);
  }
  int check_err();
};

template <>
int ocm::fill<unsigned long long>(unsigned long long* p, unsigned long long v, long long n) {
  k_ocm_fill_u64<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, v, n);
  ++launches;
  return B200ORB_OK;
}
template <>
int ocm::fill<int>(int* p, int v, long long n) {
  k_ocm_fill_i32<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, v, n);
  ++launches;
  return B200ORB_OK;
}
template <>
int ocm::fill<float>(float* p, float v, long long n) {
  k_ocm_fill_f32<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, v, n);
  ++launches;
  return B200ORB_OK;
}

static long long pow2_at_least(long long v) {
  long long p = 1;
  while (p < v) p <<= 1;
  return p;
}

int ocm::ensure_scratch(int r, int c, int nslots) {
  if (r != rows || c != cols) {
    B200_CUDA(cudaStreamSynchronize(stream));
    free_scratch();
    rows = r; cols = c;
  }
  nslots = std::min(std::max(nslots, 1), (int)MAX_SLOTS);
  if (!d_jobs) {
    B200_CUDA(cudaMalloc(&d_jobs, sizeof(KfJob) * MAX_SLOTS * JOB_RING));
    B200_CUDA(cudaHostAlloc(&h_jobs, sizeof(KfJob) * MAX_SLOTS * JOB_RING, cudaHostAllocDefault));
    for (cudaEvent_t& e : job_ev) B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const long long npix = (long long)r * c;
  const long long leaf_cap = pow2_at_least(2 * npix);
  while ((int)slots.size() < nslots) {
    KfScratch k{};
    const int id = (int)slots.size();
    k.leaf.mask = leaf_cap - 1;
    B200_CUDA(cudaMalloc(&k.leaf.rec, sizeof(LeafRec) * leaf_cap));
    B200_CUDA(cudaMalloc(&k.pix_slot, 4 * npix)); B200_CUDA(cudaMalloc(&k.pix_rank, 4 * npix)); B200_CUDA(cudaMalloc(&k.bucket, 4 * npix));
    B200_CUDA(cudaMalloc(&k.voxlist, 4 * npix));
    B200_CUDA(cudaMalloc(&k.pts, 12 * npix)); B200_CUDA(cudaMalloc(&k.pts_rgb, 3 * npix)); B200_CUDA(cudaMalloc(&k.pts_label, npix));
    k.counters = d_counters + 4 * id;
    k_ocm_fill_leaf<<<(unsigned)((leaf_cap + 255) / 256), 256, 0, stream>>>(k.leaf.rec, leaf_cap);
    ++launches;
    B200_CUDA(cudaGetLastError());
    slots.push_back(k);
  }
  return B200ORB_OK;
}

int ocm::check_err() {
  int e = 0;
  B200_CUDA(cudaMemcpyAsync(&e, d_err, 4, cudaMemcpyDeviceToHost, stream));
  B200_CUDA(cudaStreamSynchronize(stream));
  if (e) {
    B200_CUDA(cudaMemsetAsync(d_err, 0, 4, stream));
    static const char* what[] = {"", "VoxelGrid table full", "", "", "map full (raise OcmParams.map_capacity)"};
    set_error("occupancy insert: %s", what[e < 5 ? e : 0]);
    return B200ORB_ECAP;
  }
  return B200ORB_OK;
}

// n keyframes (insertion order = array order), up to MAX_SLOTS per round of launches: every stage is one launch for
// the whole round, including the map update (per-voxel replay of the round's observations in keyframe order).
int ocm::insert_batch(int n, const float* const* dd, const uint8_t* const* drgb, const uint8_t* const* dlabel, int r, int c,
                      const float* Tcw_all, float fx, float fy, float cx, float cy) {
  B200_CHECK(ensure_scratch(r, c, n));
  const int npix = r * c;
  for (int b0 = 0; b0 < n; b0 += (int)slots.size()) {
    const int B = std::min((int)slots.size(), n - b0);
    const int ring = job_head;
    job_head = (job_head + 1) % JOB_RING;
    B200_CUDA(cudaEventSynchronize(job_ev[ring]));   // the batch that used this staging entry has been uploaded
    KfJob* hj = h_jobs + (size_t)ring * MAX_SLOTS;
    KfJob* dj = d_jobs + (size_t)ring * MAX_SLOTS;
    for (int j = 0; j < B; ++j) {
      const float* Tcw = Tcw_all + 16 * (b0 + j);
      KfJob& J = hj[j];
      OcmConst& k = J.c;
      memset(&k, 0, sizeof(k));
      k.fx = fx; k.fy = fy; k.cx = cx; k.cy = cy;
      k.depth_min = prm.depth_min; k.depth_max = prm.depth_max; k.y_max = prm.y_max; k.leaf = prm.leaf;
      k.inv_leaf = prm.leaf > 0 ? 1.0f / prm.leaf : 0.f;
      k.res = prm.resolution; k.res_factor = 1.0 / prm.resolution;
      k.hit_log = hit_log; k.miss_log = miss_log; k.cmin = cmin; k.cmax = cmax;
      double R[9], t[3];
      for (int i = 0; i < 3; ++i) { for (int q = 0; q < 3; ++q) R[i * 3 + q] = (double)Tcw[i * 4 + q]; t[i] = (double)Tcw[i * 4 + 3]; }
      for (int i = 0; i < 3; ++i) {
        for (int q = 0; q < 3; ++q) k.Rt[i * 3 + q] = R[q * 3 + i];
        k.ti[i] = -(R[0 * 3 + i] * t[0] + R[1 * 3 + i] * t[1] + R[2 * 3 + i] * t[2]);
      }
      k.origin[0] = Tcw[3]; k.origin[1] = Tcw[7]; k.origin[2] = Tcw[11];
      k.rows = r; k.cols = c;
      J.depth = dd[b0 + j]; J.rgb = drgb[b0 + j]; J.label = dlabel ? dlabel[b0 + j] : nullptr;
      J.s = slots[j];
    }
    B200_CUDA(cudaMemcpyAsync(dj, hj, sizeof(KfJob) * B, cudaMemcpyHostToDevice, stream));
    B200_CUDA(cudaEventRecord(job_ev[ring], stream));
    B200_CUDA(cudaMemsetAsync(d_counters, 0, 16 * B, stream));
    const dim3 g256((npix + 255) / 256, B), g128((npix + 127) / 128, B);
    if (prm.leaf > 0) {
      const dim3 gt(tile_blocks(r, c, 8), B);   // 8 x 4 pixel tile per warp
      k_ocm_bin<<<gt, 256, 0, stream>>>(dj, d_err);
      k_ocm_ranges<<<g256, 256, 0, stream>>>(dj);
      k_ocm_scatter<<<gt, 256, 0, stream>>>(dj);
      k_ocm_centroids<<<g128, 128, 0, stream>>>(dj);
      launches += 4;
    } else {
      k_ocm_points_nofilter<<<g256, 256, 0, stream>>>(dj);
      launches += 1;
    }
    k_ocm_scan_keys<<<g256, 256, 0, stream>>>(dj, map, d_err);
    // grid-stride sweep on a fixed grid (8 CTAs per SM)
    k_ocm_apply<<<1184, 256, 0, stream>>>(hit_log, miss_log, cmin, cmax, map);
    launches += 2;
    last_slot = B - 1;
    B200_CUDA(cudaGetLastError());
  }
  return B200ORB_OK;
}

extern "C" {

void ocm_default_params(OcmParams* p) {
  if (!p) return;
  p->resolution = 0.05; p->prob_hit = 0.7; p->prob_miss = 0.4; p->clamp_min = 0.12; p->clamp_max = 0.97;
  p->depth_min = 0.5f; p->depth_max = 3.0f; p->y_max = 3.0f; p->leaf = 0.01f; p->map_capacity = 0;
}

int ocm_create(const OcmParams* p, int device, ocm_t** out) {
  if (!p || !out) { set_error("null argument"); return B200ORB_EINVAL; }
  *out = nullptr;
  if (!(p->resolution > 0) || !(p->prob_hit > 0 && p->prob_hit < 1) || !(p->prob_miss > 0 && p->prob_miss < 1) ||
      !(p->clamp_min > 0 && p->clamp_min < p->clamp_max && p->clamp_max < 1)) {
    set_error("bad OcmParams");
    return B200ORB_EINVAL;
  }
  B200_CHECK(check_device(device));
  DeviceGuard g(device);
  ocm* h = new (std::nothrow) ocm();
  if (!h) { set_error("out of host memory"); return B200ORB_EINVAL; }
  h->prm = *p;
  h->device = device;
  auto logodds = [](double pr) { return (float)log(pr / (1 - pr)); };   // octomap::logodds (float)
  h->hit_log = logodds(p->prob_hit); h->miss_log = logodds(p->prob_miss);
  h->cmin = logodds(p->clamp_min); h->cmax = logodds(p->clamp_max);
  h->map_cap = pow2_at_least(p->map_capacity > 0 ? std::max<long long>(p->map_capacity, 2) : (1ll << 23));
  auto fail = [&](cudaError_t e) { set_error("ocm_create: %s", cudaGetErrorString(e)); delete h; return B200ORB_ECUDA; };
  cudaError_t e;
  {
    // highest stream priority: the mapper is a chain of small dependent kernels running BESIDE the tracking batch;
    // with default priority every link waits behind thousands of queued tracking CTAs (measured 205 us per keyframe
    // against 74 us of kernel time), with high priority its CTAs take the next free SM slots.
    int lo_p = 0, hi_p = 0;
    cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
    if ((e = cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, hi_p)) != cudaSuccess) return fail(e);
  }
  const long long C = h->map_cap;
  h->map.mask = C - 1;
  if ((e = cudaMalloc(&h->map.keys, 8 * C)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.val, 4 * C)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.a, 4 * C)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.lo, 4 * C)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.hi, 4 * C)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.rgb, 4 * C)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.nleaves, 4)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.bm, 8 * C)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.bm_rgb, 4 * C)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.base, 4 * C)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.elist, 4 * C)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.nelist, 4)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->map.nupd, 8)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->d_counters, 16 * ocm::MAX_SLOTS)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->d_err, 4)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&h->d_export_counter, 8)) != cudaSuccess) return fail(e);
  h->fill(h->map.keys, EMPTY_KEY, C); h->fill(h->map.val, 0.f, C); h->fill(h->map.a, 0.f, C);
  h->fill(h->map.lo, -INFINITY, C); h->fill(h->map.hi, INFINITY, C);
  cudaMemsetAsync(h->map.rgb, 0xff, 4 * C, h->stream);
  cudaMemsetAsync(h->map.nleaves, 0, 4, h->stream);
  cudaMemsetAsync(h->map.bm, 0, 8 * C, h->stream);
  cudaMemsetAsync(h->map.bm_rgb, 0, 4 * C, h->stream);
  cudaMemsetAsync(h->map.base, 0, 4 * C, h->stream);
  cudaMemsetAsync(h->map.nelist, 0, 4, h->stream);
  cudaMemsetAsync(h->map.nupd, 0, 8, h->stream);
  cudaMemsetAsync(h->d_err, 0, 4, h->stream);
  if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return fail(e);
  *out = h;
  return B200ORB_OK;
}
void ocm_destroy(ocm_t* h) { delete h; }

int ocm_insert_keyframe_device(ocm_t* h, const float* d_depth, const uint8_t* d_rgb, int rows, int cols,
                               const float Tcw[16], float fx, float fy, float cx, float cy, const uint8_t* d_label) {
  if (!h || !d_depth || !d_rgb || !Tcw || rows <= 0 || cols <= 0) { set_error("bad argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  return h->insert(d_depth, d_rgb, d_label, rows, cols, Tcw, fx, fy, cx, cy);
}

int ocm_insert_keyframes_labeled_device(ocm_t* h, const float* d_depth, const uint8_t* d_rgb, const uint8_t* d_label,
                                        int rows, int cols, const int32_t* depth_idx, const int32_t* rgb_idx,
                                        const int32_t* label_idx, int n, const float* Tcw, float fx, float fy, float cx,
                                        float cy) {
  if (!h || !d_depth || !d_rgb || !depth_idx || !Tcw || rows <= 0 || cols <= 0 || n < 0) { set_error("bad argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  const size_t npix = (size_t)rows * cols;
  std::vector<const float*> dd(n);
  std::vector<const uint8_t*> dc(n), dl(n);
  for (int i = 0; i < n; ++i) {
    const int di = depth_idx[i], ri = rgb_idx ? rgb_idx[i] : depth_idx[i], li = label_idx ? label_idx[i] : depth_idx[i];
    if (di < 0 || ri < 0 || li < 0) { set_error("negative frame index"); return B200ORB_EINVAL; }
    dd[i] = d_depth + npix * di;
    dc[i] = d_rgb + npix * 3 * ri;
    dl[i] = d_label ? d_label + npix * li : nullptr;
  }
  if (n == 0) return B200ORB_OK;
  return h->insert_batch(n, dd.data(), dc.data(), d_label ? dl.data() : nullptr, rows, cols, Tcw, fx, fy, cx, cy);
}

int ocm_insert_keyframes_device(ocm_t* h, const float* d_depth, const uint8_t* d_rgb, int rows, int cols,
                                const int32_t* depth_idx, const int32_t* rgb_idx, int n, const float* Tcw, float fx,
                                float fy, float cx, float cy) {
  return ocm_insert_keyframes_labeled_device(h, d_depth, d_rgb, nullptr, rows, cols, depth_idx, rgb_idx, nullptr, n, Tcw, fx,
                                             fy, cx, cy);
}

int ocm_insert_keyframes_u16(ocm_t* h, const uint16_t* depth_u16, const uint8_t* rgb, int rows, int cols, int n,
                             float depth_factor, const float* Tcw, float fx, float fy, float cx, float cy) {
  return ocm_insert_keyframes_u16_labeled(h, depth_u16, rgb, nullptr, rows, cols, n, depth_factor, Tcw, fx, fy, cx, cy);
}

int ocm_insert_keyframes_u16_labeled(ocm_t* h, const uint16_t* depth_u16, const uint8_t* rgb, const uint8_t* label, int rows,
                                     int cols, int n, float depth_factor, const float* Tcw, float fx, float fy, float cx,
                                     float cy) {
  if (!h || !depth_u16 || !rgb || !Tcw || rows <= 0 || cols <= 0 || n < 0) { set_error("bad argument"); return B200ORB_EINVAL; }
  if (n == 0) return B200ORB_OK;
  DeviceGuard g(h->device);
  const size_t npix = (size_t)rows * cols, tot = npix * n;
  if ((npix % 4) != 0) { set_error("rows*cols must be a multiple of 4"); return B200ORB_EINVAL; }
  B200_CHECK(h->ensure_scratch(rows, cols));
  if (tot > h->kf_cap) {   // staging for n keyframes (grown on demand; reused by every later call)
    B200_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(h->d_kf_d16); cudaFree(h->d_kf_depth); cudaFree(h->d_kf_rgb); cudaFree(h->d_kf_label);
    h->d_kf_d16 = nullptr; h->d_kf_depth = nullptr; h->d_kf_rgb = nullptr; h->d_kf_label = nullptr; h->kf_cap = 0;
    B200_CUDA(cudaMalloc(&h->d_kf_d16, tot * 2)); B200_CUDA(cudaMalloc(&h->d_kf_depth, tot * 4)); B200_CUDA(cudaMalloc(&h->d_kf_rgb, tot * 3));
    B200_CUDA(cudaMalloc(&h->d_kf_label, tot));
    h->kf_cap = tot;
  }
  if (label) B200_CUDA(cudaMemcpyAsync(h->d_kf_label, label, tot, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaMemcpyAsync(h->d_kf_d16, depth_u16, tot * 2, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaMemcpyAsync(h->d_kf_rgb, rgb, tot * 3, cudaMemcpyHostToDevice, h->stream));
  k_depth_u16_to_f32<<<(unsigned)((tot / 4 + 255) / 256), 256, 0, h->stream>>>(
      reinterpret_cast<const ushort4*>(h->d_kf_d16), reinterpret_cast<float4*>(h->d_kf_depth), depth_factor, tot / 4);
  ++h->launches;
  std::vector<const float*> dd(n);
  std::vector<const uint8_t*> dc(n), dl(n);
  for (int i = 0; i < n; ++i) { dd[i] = h->d_kf_depth + npix * i; dc[i] = h->d_kf_rgb + npix * 3 * i; dl[i] = h->d_kf_label + npix * i; }
  return h->insert_batch(n, dd.data(), dc.data(), label ? dl.data() : nullptr, rows, cols, Tcw, fx, fy, cx, cy);
}

int ocm_insert_keyframe(ocm_t* h, const float* depth, const uint8_t* rgb, int rows, int cols, const float Tcw[16],
                        float fx, float fy, float cx, float cy, const uint8_t* ground_label) {
  if (!h || !depth || !rgb || !Tcw || rows <= 0 || cols <= 0) { set_error("bad argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  B200_CHECK(h->ensure_scratch(rows, cols));
  const size_t npix = (size_t)rows * cols;
  if (!h->d_depth) {
    B200_CUDA(cudaMalloc(&h->d_depth, npix * 4)); B200_CUDA(cudaMalloc(&h->d_rgb, npix * 3)); B200_CUDA(cudaMalloc(&h->d_label, npix));
  }
  B200_CUDA(cudaMemcpyAsync(h->d_depth, depth, npix * 4, cudaMemcpyHostToDevice, h->stream));
  B200_CUDA(cudaMemcpyAsync(h->d_rgb, rgb, npix * 3, cudaMemcpyHostToDevice, h->stream));
  if (ground_label) B200_CUDA(cudaMemcpyAsync(h->d_label, ground_label, npix, cudaMemcpyHostToDevice, h->stream));
  B200_CHECK(h->insert(h->d_depth, h->d_rgb, ground_label ? h->d_label : nullptr, rows, cols, Tcw, fx, fy, cx, cy));
  return h->check_err();
}

int ocm_last_points(ocm_t* h, float* xyz, uint8_t* rgb, int cap, int* n) {
  if (!h || !n) { set_error("null argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  int cnt[2] = {0, 0};
  if (h->slots.empty()) { *n = 0; return B200ORB_OK; }
  const KfScratch& ls = h->slots[h->last_slot];
  B200_CUDA(cudaMemcpyAsync(cnt, ls.counters, 8, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  *n = cnt[1];
  if (!xyz) return B200ORB_OK;
  if (cnt[1] > cap) { set_error("cap %d < %d points", cap, cnt[1]); return B200ORB_ECAP; }
  B200_CUDA(cudaMemcpyAsync(xyz, ls.pts, (size_t)12 * cnt[1], cudaMemcpyDeviceToHost, h->stream));
  if (rgb) B200_CUDA(cudaMemcpyAsync(rgb, ls.pts_rgb, (size_t)3 * cnt[1], cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return B200ORB_OK;
}

int64_t ocm_num_leaves(ocm_t* h) {
  if (!h) return -1;
  DeviceGuard g(h->device);
  int n = 0;
  if (cudaMemcpyAsync(&n, h->map.nleaves, 4, cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return -1;
  cudaStreamSynchronize(h->stream);
  return n;
}

static int export_common(ocm* h, int64_t cap, int64_t* n, unsigned short* d_k3, unsigned long long* d_k64, float* d_val,
                         unsigned char* d_rgb, float* d_a, float* d_lo, float* d_hi, int only_touched) {
  B200_CUDA(cudaMemsetAsync(h->d_export_counter, 0, 8, h->stream));
  k_ocm_export<<<(unsigned)((h->map_cap + 255) / 256), 256, 0, h->stream>>>(h->map, h->d_export_counter, cap, d_k3, d_k64,
                                                                          d_val, d_rgb, d_a, d_lo, d_hi, only_touched);
  ++h->launches;
  long long cnt = 0;
  B200_CUDA(cudaMemcpyAsync(&cnt, h->d_export_counter, 8, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  if (n) *n = cnt;
  if (cnt > cap) { set_error("export cap %lld < %lld entries", (long long)cap, cnt); return B200ORB_ECAP; }
  return B200ORB_OK;
}

int ocm_export_leaves(ocm_t* h, uint16_t* keys, float* logodds, uint8_t* rgb, int64_t cap, int64_t* n) {
  if (!h || !keys || !logodds || cap < 0) { set_error("bad argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  unsigned short* dk = nullptr; float* dv = nullptr; unsigned char* dc = nullptr;
  const int64_t c = cap > 0 ? cap : 1;
  B200_CUDA(cudaMalloc(&dk, 6 * c)); B200_CUDA(cudaMalloc(&dv, 4 * c)); B200_CUDA(cudaMalloc(&dc, 3 * c));
  int64_t cnt = 0;
  int rc = export_common(h, cap, &cnt, dk, nullptr, dv, dc, nullptr, nullptr, nullptr, 0);
  if (rc == B200ORB_OK) {
    cudaMemcpy(keys, dk, 6 * cnt, cudaMemcpyDeviceToHost);
    cudaMemcpy(logodds, dv, 4 * cnt, cudaMemcpyDeviceToHost);
    if (rgb) cudaMemcpy(rgb, dc, 3 * cnt, cudaMemcpyDeviceToHost);
  }
  if (n) *n = cnt;
  cudaFree(dk); cudaFree(dv); cudaFree(dc);
  return rc;
}

int ocm_query(ocm_t* h, const float xyz[3], float* logodds, int* found) {
  if (!h || !xyz || !logodds || !found) { set_error("null argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  int k[3];
  const double rf = 1.0 / h->prm.resolution;
  for (int i = 0; i < 3; ++i) {
    k[i] = (int)floor(rf * (double)xyz[i]) + 32768;
    if (k[i] < 0 || k[i] >= 65536) { *found = 0; *logodds = 0; return B200ORB_OK; }
  }
  const unsigned long long key = (unsigned long long)k[0] | ((unsigned long long)k[1] << 16) | ((unsigned long long)k[2] << 32);
  float* dv; int* df;
  B200_CUDA(cudaMalloc(&dv, 4)); B200_CUDA(cudaMalloc(&df, 4));
  k_ocm_query<<<1, 1, 0, h->stream>>>(h->map, key, dv, df);
  ++h->launches;
  cudaMemcpyAsync(logodds, dv, 4, cudaMemcpyDeviceToHost, h->stream);
  cudaMemcpyAsync(found, df, 4, cudaMemcpyDeviceToHost, h->stream);
  cudaStreamSynchronize(h->stream);
  cudaFree(dv); cudaFree(df);
  return B200ORB_OK;
}

int64_t ocm_summary_count(ocm_t* h) { return ocm_num_leaves(h); }

int ocm_export_summaries_device(ocm_t* h, uint64_t* d_keys, float* d_a, float* d_lo, float* d_hi, int64_t cap, int64_t* n) {
  if (!h || !d_keys || !d_a || !d_lo || !d_hi) { set_error("null argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  return export_common(h, cap, n, nullptr, (unsigned long long*)d_keys, nullptr, nullptr, d_a, d_lo, d_hi, 1);
}

int ocm_apply_summaries_device(ocm_t* h, const uint64_t* d_keys, const float* d_a, const float* d_lo, const float* d_hi,
                               int64_t n) {
  if (!h || (n > 0 && (!d_keys || !d_a || !d_lo || !d_hi))) { set_error("null argument"); return B200ORB_EINVAL; }
  if (n <= 0) return B200ORB_OK;
  DeviceGuard g(h->device);
  OcmConst k;
  memset(&k, 0, sizeof(k));
  k_ocm_apply_summaries<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(k, h->map, (const unsigned long long*)d_keys, d_a,
                                                                          d_lo, d_hi, n, h->d_err);
  ++h->launches;
  B200_CUDA(cudaGetLastError());
  return h->check_err();
}

// ---- multi-GPU map merge over NCCL (SURVEY §8(b) ocm_merge_nccl, §8(e)) ---------------------------------------------
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch already loaded when there is one), so the library
// has no link-time dependency on it and single-GPU users never touch it.
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* L = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!L) L = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (L) {
      api.lib = L;
      *(void**)&api.GetUniqueId = dlsym(L, "ncclGetUniqueId");
      *(void**)&api.CommInitRank = dlsym(L, "ncclCommInitRank");
      *(void**)&api.CommDestroy = dlsym(L, "ncclCommDestroy");
      *(void**)&api.CommCount = dlsym(L, "ncclCommCount");
      *(void**)&api.CommUserRank = dlsym(L, "ncclCommUserRank");
      *(void**)&api.AllGather = dlsym(L, "ncclAllGather");
      *(void**)&api.Broadcast = dlsym(L, "ncclBroadcast");
      *(void**)&api.GroupStart = dlsym(L, "ncclGroupStart");
      *(void**)&api.GroupEnd = dlsym(L, "ncclGroupEnd");
      *(void**)&api.GetErrorString = dlsym(L, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.Broadcast || !api.GroupStart || !api.GroupEnd ||
          !api.CommCount || !api.CommUserRank || !api.CommDestroy || !api.GetErrorString)
        api.lib = nullptr;
    }
  }
  if (!api.lib) { set_error("NCCL not available (dlopen libnccl.so.2 failed)"); return nullptr; }
  return &api;
}
#define B200_NCCL(api, expr)                                                                       \
  do {                                                                                             \
    ncclResult_t _r = (expr);                                                                      \
    if (_r != ncclSuccess) { set_error("%s -> %s", #expr, (api)->GetErrorString(_r)); return B200ORB_ECUDA; } \
  } while (0)

int ocm_nccl_unique_id(uint8_t id[128]) {
  NcclApi* N = nccl_api();
  if (!N || !id) return B200ORB_EINVAL;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId");
  ncclUniqueId u;
  B200_NCCL(N, N->GetUniqueId(&u));
  memcpy(id, &u, 128);
  return B200ORB_OK;
}
int ocm_nccl_comm_create(const uint8_t id[128], int rank, int world, int device, void** comm) {
  NcclApi* N = nccl_api();
  if (!N || !id || !comm || world < 1 || rank < 0 || rank >= world) { if (N) set_error("bad argument"); return B200ORB_EINVAL; }
  B200_CHECK(check_device(device));
  DeviceGuard g(device);
  ncclUniqueId u;
  memcpy(&u, id, 128);
  ncclComm_t c = nullptr;
  B200_NCCL(N, N->CommInitRank(&c, world, u, rank));
  *comm = (void*)c;
  return B200ORB_OK;
}
int ocm_nccl_comm_destroy(void* comm) {
  NcclApi* N = nccl_api();
  if (!N) return B200ORB_EINVAL;
  if (comm) N->CommDestroy((ncclComm_t)comm);
  return B200ORB_OK;
}

int ocm_merge_nccl(ocm_t* h, void* comm_, void* stream_, OcmMergeStats* st) {
  if (!h) { set_error("null argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  cudaStream_t S = h->stream;
  if (stream_ && (cudaStream_t)stream_ != S) {   // order after the caller's stream
    cudaEvent_t ev;
    B200_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    B200_CUDA(cudaEventRecord(ev, (cudaStream_t)stream_));
    B200_CUDA(cudaStreamWaitEvent(S, ev, 0));
    cudaEventDestroy(ev);
  }
  NcclApi* N = nullptr;
  ncclComm_t comm = (ncclComm_t)comm_;
  int world = 1, rank = 0;
  if (comm) {
    N = nccl_api();
    if (!N) return B200ORB_EINVAL;
    B200_NCCL(N, N->CommCount(comm, &world));
    B200_NCCL(N, N->CommUserRank(comm, &rank));
  }
  if (world > h->mg_world_cap) {
    if (h->mg_counts) cudaFree(h->mg_counts);
    if (h->mg_counts_h) cudaFreeHost(h->mg_counts_h);
    B200_CUDA(cudaMalloc(&h->mg_counts, 4 * world));
    B200_CUDA(cudaHostAlloc(&h->mg_counts_h, 4 * world, cudaHostAllocDefault));
    h->mg_world_cap = world;
  }
  // 1. counts: one int per rank (the only host synchronisation of the merge)
  if (comm) B200_NCCL(N, N->AllGather(h->map.nelist, h->mg_counts, 1, ncclInt32, comm, S));
  else B200_CUDA(cudaMemcpyAsync(h->mg_counts, h->map.nelist, 4, cudaMemcpyDeviceToDevice, S));
  B200_CUDA(cudaMemcpyAsync(h->mg_counts_h, h->mg_counts, 4 * world, cudaMemcpyDeviceToHost, S));
  B200_CUDA(cudaStreamSynchronize(S));
  std::vector<long long> off(world + 1, 0);   // byte offsets of the shards in the receive buffer (16-byte aligned)
  long long total = 0;
  for (int r = 0; r < world; ++r) {
    off[r + 1] = (long long)align_up_sz((size_t)(off[r] + 24ll * h->mg_counts_h[r]), 16);
    total += h->mg_counts_h[r];
  }
  const long long n_own = h->mg_counts_h[rank];
  if (24 * n_own > h->mg_send_cap) {
    if (h->mg_send) cudaFree(h->mg_send);
    h->mg_send_cap = std::max(24 * n_own * 2, 1ll << 20);
    B200_CUDA(cudaMalloc(&h->mg_send, h->mg_send_cap));
  }
  if (off[world] > h->mg_recv_cap) {
    if (h->mg_recv) cudaFree(h->mg_recv);
    if (h->mg_slots) cudaFree(h->mg_slots);
    h->mg_recv_cap = std::max(off[world] * 2, 1ll << 20);
    B200_CUDA(cudaMalloc(&h->mg_recv, h->mg_recv_cap));
    B200_CUDA(cudaMalloc(&h->mg_slots, h->mg_recv_cap / 24 * 4 + 64));
  }
  // 2. pack own summaries, roll own voxels back to the last merge
  if (n_own > 0) {
    k_merge_pack<<<(unsigned)((n_own + 255) / 256), 256, 0, S>>>(h->map, (int)n_own, merge_shard_at(h->mg_send, n_own));
    ++h->launches;
  }
  B200_CUDA(cudaMemsetAsync(h->map.nelist, 0, 4, S));
  // 3. exchange: every rank broadcasts its shard (exact sizes, one grouped NCCL call over NVLink)
  if (comm && world > 1) {
    B200_NCCL(N, N->GroupStart());
    for (int r = 0; r < world; ++r) {
      const size_t bytes = 24 * (size_t)h->mg_counts_h[r];
      if (bytes == 0) continue;
      B200_NCCL(N, N->Broadcast(r == rank ? h->mg_send : nullptr, (char*)h->mg_recv + off[r], bytes, ncclUint8, r, comm, S));
    }
    B200_NCCL(N, N->GroupEnd());
  } else if (n_own > 0) {
    B200_CUDA(cudaMemcpyAsync(h->mg_recv, h->mg_send, 24 * n_own, cudaMemcpyDeviceToDevice, S));
  }
  // 4. replay all shards in rank (= keyframe) order, then make the result the base of the next epoch
  long long done = 0;
  for (int r = 0; r < world; ++r) {
    const long long n = h->mg_counts_h[r];
    if (n == 0) continue;
    k_merge_apply<<<(unsigned)((n + 255) / 256), 256, 0, S>>>(h->map, n, merge_shard_at((char*)h->mg_recv + off[r], n),
                                                             h->mg_slots + done, h->d_err);
    ++h->launches;
    done += n;
  }
  if (total > 0) {
    k_merge_commit<<<(unsigned)((total + 255) / 256), 256, 0, S>>>(h->map, total, h->mg_slots);
    ++h->launches;
  }
  B200_CUDA(cudaGetLastError());
  if (stream_ && (cudaStream_t)stream_ != S) {
    cudaEvent_t ev;
    B200_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    B200_CUDA(cudaEventRecord(ev, S));
    B200_CUDA(cudaStreamWaitEvent((cudaStream_t)stream_, ev, 0));
    cudaEventDestroy(ev);
  }
  if (st) { st->world = world; st->rank = rank; st->records_sent = n_own; st->records_total = total; st->bytes_sent = 24 * n_own; st->bytes_received = 24 * (total - n_own); }
  return B200ORB_OK;
}

int ocm_last_batch_stats(ocm_t* h, int64_t* points, int64_t* voxels_touched) {
  if (!h) { set_error("null argument"); return B200ORB_EINVAL; }
  DeviceGuard g(h->device);
  std::vector<int> c(4 * ocm::MAX_SLOTS, 0);
  unsigned long long nt = 0;
  B200_CUDA(cudaMemcpyAsync(c.data(), h->d_counters, 16 * ocm::MAX_SLOTS, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaMemcpyAsync(&nt, h->map.nupd, 8, cudaMemcpyDeviceToHost, h->stream));
  B200_CUDA(cudaStreamSynchronize(h->stream));
  long long p = 0;
  for (int j = 0; j <= h->last_slot && j < ocm::MAX_SLOTS; ++j) p += c[4 * j + 1];
  if (points) *points = p;
  if (voxels_touched) *voxels_touched = (int64_t)nt;
  return B200ORB_OK;
}

int ocm_sync(ocm_t* h) {
  if (!h) return B200ORB_EINVAL;
  DeviceGuard g(h->device);
  B200_CUDA(cudaStreamSynchronize(h->stream));
  return h->check_err();
}
void* ocm_stream(ocm_t* h) { return h ? (void*)h->stream : nullptr; }
long long ocm_launch_count(const ocm_t* h) { return h ? h->launches : 0; }

}  // extern "C"
