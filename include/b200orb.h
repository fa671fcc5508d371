/* b200orb.h -- C-ABI of libb200orb.so: the B200 (sm_100a) implementation of the ORB-SLAM2 hot path.
 *
 * The reference (Ewenwan/ORB_SLAM2_SSD_Semantic) has no FFI layer; its boundary is the public surface of
 * three C++ classes.  Every entry point below names the reference member it replaces
 * (file:line relative to the reference root).  INTEGRATION.md shows the header-only C++ shim classes
 * (orb_slam2_ssd_semantic_b200/csrc/shim/) that keep those class surfaces source-compatible on top of this ABI.
 *
 * Conventions: plain C, caller-owned buffers, int status (0 = OK, <0 = B200ORB_E*), never aborts/exits,
 * no exceptions cross the boundary.  Opaque handles own device memory and one CUDA stream each; calls on
 * DISTINCT handles are re-entrant (the reference runs L/R extractors on two threads, src/Frame.cc:121-124,
 * and matchers on three).  Pointers named d_* are CUDA device pointers, everything else is host memory.
 */
#ifndef B200ORB_H_
#define B200ORB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ORB_OK 0
#define B200ORB_EINVAL (-1)     /* bad argument */
#define B200ORB_ECAP (-2)       /* caller buffer too small */
#define B200ORB_ECUDA (-3)      /* CUDA runtime error (see b200orb_last_error) */
#define B200ORB_ENOGPU (-4)     /* no usable CUDA device: there is NO CPU fallback */
#define B200ORB_EGEOM (-5)      /* image geometry the reference itself cannot process */

const char* b200orb_last_error(void);   /* thread-local message of the last failing call */
int b200orb_device_count(void);         /* number of CUDA devices visible (0 if none / no driver) */
const char* b200orb_version(void);
/* Bit mask of the round-2 kernel formulations in force (default 3; environment B200ORB_EXPERIMENTAL overrides it, 0 = the
 * first formulations, kept for A/B): bit 0 k_orient_desc2, bit 1 second FAST tile staging.  Results never depend on it. */
int b200orb_experimental(void);
/* Launch-shape tuning of the extractor (process-wide; results never depend on it): FAST cells per CTA (1, 2, 4, 8) and the
 * quad-tree's register budget (2, 3 or 4 CTAs per SM); exp_mask as above.  Negative arguments keep the current value.
 * Environment at load: B200ORB_EXPERIMENTAL, B200ORB_FAST_WPC, B200ORB_QT_MINB. */
int b200orb_get_tuning(int* exp_mask, int* fast_wpc, int* qt_minblocks);
int b200orb_set_tuning(int exp_mask, int fast_wpc, int qt_minblocks);

/* ------------------------------------------------------------------------------------------------
 * ORB extractor  --  ORB_SLAM2::ORBextractor (include/ORBextractor.h:41-118, src/ORBextractor.cc)
 * ------------------------------------------------------------------------------------------------ */
typedef struct orbx orbx_t;

typedef struct {            /* ctor arguments, src/ORBextractor.cc:399-400 */
  int nfeatures;
  float scale_factor;
  int nlevels;
  int ini_th_fast;
  int min_th_fast;
} OrbxParams;

typedef struct {            /* cv::KeyPoint, 28 bytes, what operator() appends to _keypoints */
  float x, y, size, angle, response;
  int32_t octave, class_id;
} OrbxKeyPoint;

/* ORBextractor::ORBextractor (src/ORBextractor.cc:399-466).  device = CUDA ordinal. */
int orbx_create(const OrbxParams* p, int device, orbx_t** out);
void orbx_destroy(orbx_t* h);

/* Upper bound on keypoints one frame can return: nfeatures + 3*nlevels. */
int orbx_max_keypoints(const orbx_t* h);

/* ORBextractor::operator() (src/ORBextractor.cc:1052-1114) on one HOST image (CV_8UC1, row stride in bytes).
 * kps[cap], desc[cap*32] are host buffers; *n_out keypoints are written level-major exactly as the
 * reference orders them.  rows==0||cols==0 -> *n_out = 0, OK (reference: silent return, :1055). */
int orbx_extract(orbx_t* h, const uint8_t* gray, int rows, int cols, size_t stride,
                 OrbxKeyPoint* kps, uint8_t* desc, int cap, int* n_out);

/* Batched many-frame mode (north_star): nframes images of identical geometry, frame f at
 * gray + f*frame_stride.  Outputs: kps[f*cap ..], desc[(f*cap)*32 ..], n_out[f].  Host buffers. */
int orbx_extract_batch(orbx_t* h, const uint8_t* gray, int nframes, int rows, int cols, size_t stride,
                       size_t frame_stride, OrbxKeyPoint* kps, uint8_t* desc, int cap, int* n_out);

/* Same, with the images already resident in HBM (d_gray) and results LEFT in HBM: after the call
 * orbx_device_results() returns device pointers valid until the next extract on this handle.
 * Asynchronous on the handle's stream; orbx_sync() waits. */
int orbx_extract_batch_device(orbx_t* h, const uint8_t* d_gray, int nframes, int rows, int cols,
                              size_t stride, size_t frame_stride);
int orbx_device_results(orbx_t* h, const OrbxKeyPoint** d_kps, const uint8_t** d_desc,
                        const int32_t** d_counts, int* cap);
int orbx_sync(orbx_t* h);
void* orbx_stream(orbx_t* h);   /* cudaStream_t the handle launches on */

/* Backs the public member mvImagePyramid (include/ORBextractor.h:80): level image of frame `frame` of the
 * last extract.  bordered=0: rows x cols level (the ROI the reference stores); bordered=1: the
 * (rows+38)x(cols+38) parent buffer with the BORDER_REFLECT_101 frame (src/ORBextractor.cc:1125-1142). */
int orbx_level_dims(const orbx_t* h, int level, int* rows, int* cols);
int orbx_get_level(orbx_t* h, int frame, int level, int bordered, uint8_t* dst, size_t dst_stride);

/* GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:63-77) + mnFeaturesPerLevel; each array has nlevels entries (NULL = skip). */
int orbx_scale_tables(const orbx_t* h, float* sf, float* inv_sf, float* sigma2, float* inv_sigma2,
                      int* features_per_level);

/* Diagnostics for parity tests: FAST candidates per level of frame `frame` of the last extract. */
int orbx_candidates_per_level(orbx_t* h, int frame, int* counts /*nlevels*/);
/* Number of CUDA kernels launched by this handle since creation (bench.py's gpu_launches). */
long long orbx_launch_count(const orbx_t* h);
/* Per-stage device time via CUDA events on the handle's stream (bench.py's roofline leg).  Stages:
 * 0 pyramid resize, 1 FAST cells, 2 quad-tree, 3 blur, 4 orientation+descriptor, 5 frame glue, 6 match
 * (5,6 only when driven through orbs_*).  read() sums ms per stage over all runs since the last read. */
#define B200ORB_NUM_STAGES 7
int orbx_profile_enable(orbx_t* h, int on);
int orbx_profile_read(orbx_t* h, float* ms, long long* frames, int* runs);

/* ------------------------------------------------------------------------------------------------
 * ORB matcher  --  ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:37-118, src/ORBmatcher.cc)
 * ------------------------------------------------------------------------------------------------ */
#define ORBM_TH_HIGH 100      /* src/ORBmatcher.cc:39 */
#define ORBM_TH_LOW 50        /* :40 */
#define ORBM_HISTO_LENGTH 30  /* :41 */
#define ORBM_GRID_COLS 64     /* include/Frame.h:26 */
#define ORBM_GRID_ROWS 48     /* include/Frame.h:25 */

/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1968-1984): Hamming distance of two 256-bit rows. */
int orbm_hamming(const uint8_t a[32], const uint8_t b[32]);

typedef struct orbm orbm_t;   /* workspace + stream; the matcher itself is stateless like the reference's */
int orbm_create(int device, orbm_t** out);
void orbm_destroy(orbm_t* h);
long long orbm_launch_count(const orbm_t* h);

/* Flat view of the "current" Frame the searches read (src/Frame.cc / include/Frame.h). */
typedef struct {
  int n;                      /* Frame::N */
  const float* x;             /* mvKeysUn[i].pt.x */
  const float* y;             /* mvKeysUn[i].pt.y */
  const int32_t* octave;      /* mvKeysUn[i].octave */
  const float* angle;         /* mvKeysUn[i].angle */
  const float* uright;        /* mvuRight[i] (<=0: none) */
  const uint8_t* desc;        /* mDescriptors, n x 32 */
  const int32_t* mp_obs;      /* per keypoint: -1 if mvpMapPoints[i]==NULL else that MapPoint's Observations();
                                 NULL = all -1 (what Tracking does before the call, src/Tracking.cc:1337) */
  float Tcw[16];              /* mTcw row-major 4x4 */
  float fx, fy, cx, cy, bf, b;        /* mbf, mb */
  float min_x, max_x, min_y, max_y;   /* mnMinX.. (static image bounds) */
  const float* scale_factors; /* mvScaleFactors, nlevels entries */
  int nlevels;
} OrbmFrame;

/* "Last frame" side of SearchByProjection(Frame&, const Frame&, th, bMono). */
typedef struct {
  int n;                      /* LastFrame.N */
  const float* xw;            /* n x 3: pMP->GetWorldPos() of LastFrame.mvpMapPoints[i] (ignored where !valid) */
  const uint8_t* valid;       /* 1 iff mvpMapPoints[i] != NULL && !mvbOutlier[i] */
  const int32_t* octave;      /* LastFrame.mvKeys[i].octave */
  const float* angle;         /* LastFrame.mvKeysUn[i].angle */
  const uint8_t* mp_desc;     /* n x 32: pMP->GetDescriptor() */
  const int32_t* mp_obs;      /* pMP->Observations() (decides whether a claimed keypoint blocks later queries,
                                 src/ORBmatcher.cc:1656-1658); NULL = all 0 */
  float Tcw[16];              /* LastFrame.mTcw */
} OrbmLast;

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono)
 * (src/ORBmatcher.cc:1578-1724).  cur2last[cur.n]: index i of the LastFrame keypoint whose MapPoint ends up
 * in CurrentFrame.mvpMapPoints[j], or -1 (NULL) -- i.e. exactly the pointer state the reference leaves
 * (pre-existing entries given through cur->mp_obs are reported as -2 when they survive untouched).
 * *nmatches = the function's return value (may double-count).  Host buffers. */
int orbm_search_by_projection_last(orbm_t* h, const OrbmFrame* cur, const OrbmLast* last, float th, int mono,
                                   float nnratio, int check_ori, int32_t* cur2last, int* nmatches);

/* Local-map search: ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*>&, th)
 * (src/ORBmatcher.cc:63-156); the per-MapPoint fields are those Frame::isInFrustum fills (src/Frame.cc:387-451). */
typedef struct {
  int n;
  const uint8_t* track_in_view;   /* mbTrackInView && !isBad() */
  const float* proj_x;            /* mTrackProjX */
  const float* proj_y;            /* mTrackProjY */
  const float* proj_xr;           /* mTrackProjXR */
  const int32_t* scale_level;     /* mnTrackScaleLevel */
  const float* view_cos;          /* mTrackViewCos */
  const uint8_t* mp_desc;         /* n x 32 */
  const int32_t* mp_obs;          /* Observations(); NULL = all 1 */
} OrbmTrackPoints;
int orbm_search_by_projection_points(orbm_t* h, const OrbmFrame* f, const OrbmTrackPoints* pts, float th,
                                     float nnratio, int32_t* f2pt, int* nmatches);

/* Generic guided search for the remaining projection overloads: the caller (shim) has already projected its MapPoints
 * and chosen window + octave range per query -- SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist)
 * (src/ORBmatcher.cc:1757-1899, relocalisation) and the Sim3 / loop-closing variants do that part in a dozen
 * scalar lines each (PredictScale, distance gates) -- and the GPU does what they share: GetFeaturesInArea, the Hamming
 * loop, the ordered "already matched" rule, the rotation histogram.
 *   claim_rule 0: a keypoint blocks later queries iff the MapPoint holding it has Observations()>0 (:1656-1658)
 *   claim_rule 1: any MapPoint already on the keypoint blocks it (:1830-1831 `if(CurrentFrame.mvpMapPoints[i2]) continue`)
 * cur2q[j] = query index now on keypoint j (-1 NULL, -2 untouched pre-existing). */
typedef struct {
  int n;
  const uint8_t* valid;        /* query takes part (MapPoint non-NULL, !isBad(), in image, distance gates passed) */
  const float* u;              /* projected pixel */
  const float* v;
  const float* radius;         /* th * mvScaleFactors[nPredictedLevel] */
  const int32_t* min_level;    /* GetFeaturesInArea(minLevel, maxLevel) semantics (src/Frame.cc:486-506) */
  const int32_t* max_level;
  const float* uright;         /* predicted right coordinate for the stereo gate; NULL = no gate */
  const uint8_t* desc;         /* n x 32: pMP->GetDescriptor() */
  const float* angle;          /* angle used for the rotation histogram (e.g. pKF->mvKeysUn[i].angle) */
  const int32_t* obs;          /* Observations() per query (claim_rule 0); NULL = all 1 */
} OrbmQueries;
int orbm_search_projected(orbm_t* h, const OrbmFrame* cur, const OrbmQueries* q, int max_dist, int claim_rule,
                          int check_ori, int32_t* cur2q, int* nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:217-363).
 * The two DBoW2::FeatureVector maps are given flattened and sorted by node id:
 * node_ids[k] (strictly ascending), indices of node k = idx[node_off[k] .. node_off[k+1]). */
typedef struct {
  int n;                       /* keypoints */
  const uint8_t* desc;         /* n x 32 */
  const float* angle;          /* KF: mvKeysUn[i].angle ; F: mvKeys[i].angle (src/ORBmatcher.cc:301,308) */
  const uint8_t* valid;        /* KF side: MapPoint non-NULL && !isBad(); F side: NULL */
  int n_nodes;
  const uint32_t* node_ids;
  const int32_t* node_off;     /* n_nodes+1 */
  const uint32_t* idx;
} OrbmBow;
int orbm_search_by_bow(orbm_t* h, const OrbmBow* kf, const OrbmBow* f, float nnratio, int check_ori,
                       int32_t* f2kf /* f->n, -1 = none */, int* nmatches);
/* ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) (src/ORBmatcher.cc:665-812,
 * loop closing): both sides carry `valid` (MapPoint non-NULL && !isBad()); matches12[kf1->n] = index of the pKF2
 * keypoint whose MapPoint lands in vpMatches12[i], or -1. */
int orbm_search_by_bow_kf(orbm_t* h, const OrbmBow* kf1, const OrbmBow* kf2, float nnratio, int check_ori,
                          int32_t* matches12, int* nmatches);

/* Per-query best keypoint of a keyframe, with NO claim state: the device part of
 *   Fuse(KeyFrame*, const vector<MapPoint*>&, th)            src/ORBmatcher.cc:1031-1182   gate 1 (reprojection chi-square,
 *                                                            :1118-1146: needs q->uright = u - bf*invz and inv_level_sigma2)
 *   Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)       src/ORBmatcher.cc:1198-1318   gate 0
 *   SearchBySim3 (each direction)                            src/ORBmatcher.cc:1334-1558   gate 0
 * The caller (shim) projects its MapPoints and passes the distance / viewing-angle gates, PredictScale and the radius
 * exactly as those loop heads do (queries: u, v, radius, min_level = nPredictedLevel-1, max_level = nPredictedLevel,
 * descriptor); the device walks KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:659-698), applies the level range and the
 * gate and returns the FIRST minimum of the Hamming distance: best_idx[i] (-1: no candidate), best_dist[i].  What happens
 * with a match (Replace / AddObservation / mutual-consistency test) stays with the caller, in the reference's order --
 * none of it feeds back into later queries' candidate sets. */
int orbm_search_best(orbm_t* h, const OrbmFrame* kf, const OrbmQueries* q, int gate, const float* inv_level_sigma2,
                     int32_t* best_idx, int32_t* best_dist);

/* ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<Point2f>& vbPrevMatched, vector<int>& vnMatches12,
 * int windowSize) (src/ORBmatcher.cc:523-651).  F1 supplies octave / angle / descriptors, F2 additionally its grid;
 * prev_xy[n1][2] = vbPrevMatched (in / out); matches12[n1] = vnMatches12. */
int orbm_search_for_initialization(orbm_t* h, const OrbmFrame* f1, const OrbmFrame* f2, float* prev_xy, int window,
                                   float nnratio, int check_ori, int32_t* matches12, int* nmatches);

/* ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vector<pair<size_t,size_t>>&,
 * bOnlyStereo) (src/ORBmatcher.cc:827-1019).  Both keyframes flattened with their FeatureVector (as OrbmBow), undistorted
 * keypoints, right coordinates and MapPoint occupancy; F12 row-major 3x3; (ex, ey) = the epipole the caller computes from
 * the two poses (:835-839); scale_factors2 / level_sigma2_2 = pKF2->mvScaleFactors / mvLevelSigma2.  matches12[k1->n] =
 * pKF2 keypoint matched to pKF1 keypoint i or -1 (the caller turns it into vMatchedPairs in index order). */
typedef struct {
  int n;
  const uint8_t* desc;         /* n x 32 */
  const float* x;              /* mvKeysUn[i].pt */
  const float* y;
  const float* angle;
  const float* uright;         /* mvuRight (>= 0: stereo) */
  const int32_t* octave;
  const uint8_t* has_mp;       /* GetMapPoint(i) != NULL */
  int n_nodes;
  const uint32_t* node_ids;
  const int32_t* node_off;
  const uint32_t* idx;
} OrbmTriKF;
int orbm_search_for_triangulation(orbm_t* h, const OrbmTriKF* kf1, const OrbmTriKF* kf2, const float F12[9], float ex,
                                  float ey, const float* scale_factors2, const float* level_sigma2_2, int nlevels,
                                  int only_stereo, int check_ori, int32_t* matches12, int* nmatches);

/* Frame::isInFrustum (src/Frame.cc:387-451) for n MapPoints at once -- what Tracking::SearchLocalPoints runs per local
 * MapPoint before SearchByProjection(Frame&, vector<MapPoint*>&, th) (src/Tracking.cc:1931-1946): projection, image and
 * distance-invariance gates, viewing-angle gate, MapPoint::PredictScale.  The outputs are exactly the per-MapPoint fields
 * OrbmTrackPoints wants (mbTrackInView, mTrackProjX/Y/XR, mnTrackScaleLevel, mTrackViewCos); entries not in view are 0.
 * log_scale_factor = Frame::mfLogScaleFactor; min_dist / max_dist = mfMinDistance / mfMaxDistance (raw members). */
typedef struct {
  int n;
  const float* xw;        /* n x 3 GetWorldPos() */
  const float* normal;    /* n x 3 GetNormal() */
  const float* min_dist;  /* mfMinDistance */
  const float* max_dist;  /* mfMaxDistance */
} OrbmFrustumPoints;
int orbm_is_in_frustum(orbm_t* h, const OrbmFrame* f, const OrbmFrustumPoints* p, float viewing_cos_limit,
                       float log_scale_factor, uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr,
                       int32_t* scale_level, float* view_cos);

/* Frame::UndistortKeyPoints (src/Frame.cc:559-590): cv::undistortPoints(pts, pts, mK, mDistCoef, Mat(), mK) on n keypoint
 * positions (xy interleaved); K row-major 3x3, dist = k1 k2 p1 p2 [k3].  dist[0] == 0 copies the input (:561-565). */
int orbm_undistort_keypoints(orbm_t* h, const float* xy_in, int n, const float K[9], const float* dist, int ndist,
                             float* xy_out);

/* ------------------------------------------------------------------------------------------------
 * BoW transform -- Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:546-555, src/KeyFrame.cc:75-84):
 * mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4) of DBoW2's TemplatedVocabulary (k-ary tree; ORBvoc: k = 10,
 * L = 6).  orbv_create uploads the tree: node 0 = root, parent[i] < i, children visited in ascending node id, leaves are
 * the words numbered in ascending node id, desc n_nodes x 32, weight = node weights (TF-IDF: the idf of the word).
 * orbv_transform descends every descriptor (first child at minimal Hamming distance per level) and returns per feature
 * the word id, the word's weight and the id of the ancestor `levelsup` levels above the leaves (0 = root when L - levelsup
 * <= 0).  The caller builds the maps in feature order exactly like TemplatedVocabulary::transform does:
 *   if (weight > 0) { BowVector.addWeight(word, weight); FeatureVector.addFeature(node, i); }  then  BowVector.normalize(L1).
 * ------------------------------------------------------------------------------------------------ */
typedef struct orbv orbv_t;
int orbv_create(int device, int k, int L, int n_nodes, const int32_t* parent, const uint8_t* desc, const double* weight,
                const int32_t* word_id /* per node, read at leaves; NULL = leaves numbered in ascending node id */, orbv_t** out);
void orbv_destroy(orbv_t* h);
int orbv_num_words(const orbv_t* h);
long long orbv_launch_count(const orbv_t* h);
int orbv_transform(orbv_t* h, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* word_weight,
                   uint32_t* node_id);

/* ------------------------------------------------------------------------------------------------
 * Dynamic-mask stages of the `perfect` variant (the image-arithmetic parts):
 *   FlowSLAM::Flow::ComputeMask (perfect/include/Flow.h:24-27, perfect/src/Flow.cc:17-52) after the call of
 *   cv::calcOpticalFlowFarneback -- which stays on the host side: OpenCV's own float code -- i.e. pyrUp of the half-
 *   resolution flow field (:30), mask = 1 except 0 where x*x + y*y >= max(BInaryThreshold, 40) (:24,31-41), then
 *   erode, erode, dilate with getStructuringElement(MORPH_ELLIPSE, 21x21) (:42-47);
 *   and the keypoint loop of the masked RGB-D Frame constructor (perfect/src/Frame.cc:356-377): if cv::sum(mask) >
 *   rows*cols*0.65, keypoints whose pixel mask.at<uchar>(pt.y, pt.x) is not 1 are dropped together with their
 *   descriptor rows, order kept; otherwise nothing is dropped.
 * flow: rows x cols x 2 float (CV_32FC2, continuous) at half resolution; mask: mask_rows x mask_cols bytes (0 / 1),
 * continuous, the size of the gray image (2 rows <= mask_rows <= 2 rows + 1, likewise the columns: pyrDown halves with
 * truncation); pixels beyond the 2 rows x 2 cols the up-sampled flow covers keep their initial 1 (:25).
 * ------------------------------------------------------------------------------------------------ */
typedef struct dynm dynm_t;
int dynm_create(int device, dynm_t** out);
void dynm_destroy(dynm_t* h);
long long dynm_launch_count(const dynm_t* h);
void* dynm_stream(dynm_t* h);
int dynm_sync(dynm_t* h);
int dynm_element(const dynm_t* h, uint8_t* element /* 21 x 21: the structuring element in use */);
int dynm_mask_from_flow(dynm_t* h, const float* flow, int rows, int cols, float binary_threshold, uint8_t* mask, int mask_rows,
                        int mask_cols);
/* nframes flow fields / masks back to back in HBM, asynchronous on the handle's stream */
int dynm_mask_from_flow_batch_device(dynm_t* h, const float* d_flow, int nframes, int rows, int cols, float binary_threshold,
                                     uint8_t* d_mask, int mask_rows, int mask_cols);
/* host buffers, filtered in place; *n_out keypoints remain */
int dynm_filter_keypoints(dynm_t* h, const uint8_t* mask, int rows, int cols, size_t stride, OrbxKeyPoint* kps, uint8_t* desc,
                          int n, int* n_out);
/* the layout of orbx_device_results / orbs_device_results: kps [nframes][cap], desc [nframes][cap][32], counts [nframes],
 * masks [nframes][rows][cols]; filtered in place, asynchronous on the handle's stream */
int dynm_filter_keypoints_batch_device(dynm_t* h, const uint8_t* d_mask, int nframes, int rows, int cols, OrbxKeyPoint* d_kps,
                                       uint8_t* d_desc, int32_t* d_counts, int cap);

/* ------------------------------------------------------------------------------------------------
 * Stream pipeline (batched many-frame mode of north_star): per frame t of a batch, what
 * Tracking::GrabImageRGBD -> Frame::Frame(RGB-D) -> TrackWithMotionModel's SearchByProjection do
 * (src/Tracking.cc:331-375,1324-1352; src/Frame.cc:176-240): extract, ComputeStereoFromRGBD
 * (src/Frame.cc:850-871), AssignFeaturesToGrid (:319-334), UnprojectStereo (:879-899) of frame t-1's
 * keypoints into "map points", then SearchByProjection(frame t, frame t-1, th, mono=0).
 * ------------------------------------------------------------------------------------------------ */
typedef struct orbs orbs_t;
typedef struct {
  OrbxParams orb;
  float fx, fy, cx, cy, bf;     /* Camera.* of the YAML (perfect/Examples/RGB-D/TUM3.yaml:8-25) */
  float th;                     /* search window, 15 in src/Tracking.cc:1346 */
  float nnratio;                /* 0.9 (src/Tracking.cc:1327) */
  int check_ori;                /* 1 */
  int max_frames;               /* batch capacity */
} OrbsParams;
int orbs_create(const OrbsParams* p, int device, orbs_t** out);
void orbs_destroy(orbs_t* h);
/* Host-buffer entry: gray[n][rows][cols] u8, depth[n][rows][cols] f32 metres, Tcw[n][16] f32 row-major.
 * Outputs (host): kps[n][cap], desc[n][cap][32], nkp[n], cur2last[n][cap] (frame 0: all -1), nmatch[n]. */
int orbs_track_batch(orbs_t* h, const uint8_t* gray, const float* depth, const float* Tcw, int nframes,
                     int rows, int cols, OrbxKeyPoint* kps, uint8_t* desc, int32_t* nkp, int32_t* cur2last,
                     int32_t* nmatch, int cap);
/* Same with the depth image as the sensor delivers it (CV_16U, e.g. TUM PNGs): the conversion
 * imDepth.convertTo(CV_32F, mDepthMapFactor) of Tracking::GrabImageRGBD (src/Tracking.cc:366-367; float multiply by
 * depth_factor = 1.0f / DepthMapFactor) runs on the device, halving the host->device traffic of the depth stream. */
int orbs_track_batch_u16(orbs_t* h, const uint8_t* gray, const uint16_t* depth_u16, float depth_factor, const float* Tcw,
                         int nframes, int rows, int cols, OrbxKeyPoint* kps, uint8_t* desc, int32_t* nkp,
                         int32_t* cur2last, int32_t* nmatch, int cap);
/* When depth_u16 is page-locked host memory (cudaHostAlloc / cudaHostRegister) orbs_track_batch_u16 does not upload
 * the depth images at all: the tracking path reads depth only under the keypoints (Frame::ComputeStereoFromRGBD,
 * src/Frame.cc:850-871), so the device gathers those pixels in place over PCIe.  orbs_set_full_depth_upload(h, 1)
 * forces the full upload (then orbs_device_inputs returns the converted f32 batch). */
int orbs_set_full_depth_upload(orbs_t* h, int on);
/* Pipelining aid for two handles used alternately: the next batch submitted to h starts its kernels only when the kernels
 * of prev's last submitted batch are done; h's uploads and prev's downloads proceed meanwhile, so the stages upload(k+1) |
 * kernels(k) | download(k-1) overlap without two batches' kernels sharing the SMs.  The dependency is consumed by h's next
 * submit; prev must stay alive until then. */
int orbs_chain_after(orbs_t* h, orbs_t* prev);
/* Frames per upload chunk of the host-buffer entries (default 128, at most 7 chunks per call): the extraction of a chunk
 * starts as soon as it has arrived while the next one is still crossing PCIe. */
int orbs_set_chunk_frames(orbs_t* h, int frames);
/* Streaming form of orbs_track_batch_u16: enqueues uploads, kernels and result downloads on the handle's streams and
 * returns; the outputs are valid after orbs_sync(h).  All host buffers must be page-locked and stay untouched until
 * then.  Two handles used alternately keep two batches in flight, so the upload of batch k+1 and the download of
 * batch k-1 overlap the kernels of batch k (what the reference's grab thread / tracking thread split does for one
 * frame, Examples/RGB-D/rgbd_tum.cc:88-109, done here for whole batches). */
int orbs_submit_batch_u16(orbs_t* h, const uint8_t* gray, const uint16_t* depth_u16, float depth_factor, const float* Tcw,
                          int nframes, int rows, int cols, OrbxKeyPoint* kps, uint8_t* desc, int32_t* nkp,
                          int32_t* cur2last, int32_t* nmatch, int cap);
/* convertTo(CV_32F, factor) of n CV_16U pixels resident in HBM (n multiple of 4) on the given cudaStream_t. */
int b200orb_depth_u16_to_f32_device(const uint16_t* d_src, float* d_dst, size_t n, float factor, void* stream);
/* Device copies of the inputs of the last host-buffer call (gray u8, depth f32 metres), e.g. to hand keyframes to
 * ocm_insert_keyframes_device without a second upload. */
int orbs_device_inputs(orbs_t* h, const uint8_t** d_gray, const float** d_depth);
/* Device-resident entry (inputs already in HBM, results left in HBM; asynchronous on orbs_stream). */
int orbs_track_batch_device(orbs_t* h, const uint8_t* d_gray, const float* d_depth, const float* d_Tcw,
                            int nframes, int rows, int cols);
int orbs_device_results(orbs_t* h, const OrbxKeyPoint** d_kps, const uint8_t** d_desc, const int32_t** d_nkp,
                        const int32_t** d_cur2last, const int32_t** d_nmatch, int* cap);
/* Frame glue of frame `frame` of the last batch: mvuRight / mvDepth (Frame::ComputeStereoFromRGBD, src/Frame.cc:850-871), the
 * world point Frame::UnprojectStereo gives per keypoint under that frame's pose (:879-899) and whether the keypoint has
 * depth; host arrays of at least the handle's keypoint capacity (NULL = skip). */
int orbs_read_frame_glue(orbs_t* h, int frame, float* uright, float* depth, float* xw, uint8_t* valid, int cap);
int orbs_sync(orbs_t* h);
void* orbs_stream(orbs_t* h);
long long orbs_launch_count(const orbs_t* h);
orbx_t* orbs_extractor(orbs_t* h);

/* ------------------------------------------------------------------------------------------------
 * Dense mapping -- PointCloudMapping (include/pointcloudmapping.h:50-56) with the OctoMap occupancy
 * semantics of perfect/src/MapDrawer.cc:51-56,610-675,946-1025.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ocm ocm_t;
typedef struct {
  double resolution;            /* octoMap.res, 0.05 (perfect/Examples/RGB-D/my_rgbd_ty_api_adj.yaml:82) */
  double prob_hit, prob_miss;   /* 0.7 / 0.4   (perfect/src/MapDrawer.cc:55-56) */
  double clamp_min, clamp_max;  /* 0.12 / 0.97 (perfect/src/MapDrawer.cc:53-54) */
  float depth_min, depth_max;   /* 0.5 / 3.0   (perfect/src/MapDrawer.cc:655) */
  float y_max;                  /* 3.0: keep |y| <= y_max (:660) */
  float leaf;                   /* 0.01 VoxelGrid leaf (:669); <=0 disables the pre-filter */
  int64_t map_capacity;         /* voxel hash capacity (slots); 0 = default */
} OcmParams;
void ocm_default_params(OcmParams* p);
int ocm_create(const OcmParams* p, int device, ocm_t** out);
void ocm_destroy(ocm_t* h);
/* One keyframe: MapDrawer::GeneratePointCloud (:641-675) + InsertScan (:946-1025) with a supplied
 * ground label per pixel (NULL = all non-ground, the reference's fallback when RANSAC finds no plane).
 * depth f32 metres, rgb u8 BGR-interleaved as cv::Mat (kf->mImRGB), Tcw row-major (KeyFrame::GetPose()). */
int ocm_insert_keyframe(ocm_t* h, const float* depth, const uint8_t* rgb, int rows, int cols,
                        const float Tcw[16], float fx, float fy, float cx, float cy,
                        const uint8_t* ground_label);
int ocm_insert_keyframe_device(ocm_t* h, const float* d_depth, const uint8_t* d_rgb, int rows, int cols,
                               const float Tcw[16], float fx, float fy, float cx, float cy,
                               const uint8_t* d_ground_label);
/* Batched variant for a resident RGB-D stream: inserts keyframes frame_idx[0..n) (in that order) of a batch laid out
 * as depth[F][rows][cols] f32, rgb[F][rows][cols][3] u8 in HBM, Tcw[n][16] on the host (one pose per inserted
 * keyframe); keyframe i reads depth image depth_idx[i] and colour image rgb_idx[i] of the two batches.  One call
 * enqueues everything on the handle's stream (what UpdateOctomap's loop does,
 * perfect/src/MapDrawer.cc:610-638). */
int ocm_insert_keyframes_device(ocm_t* h, const float* d_depth, const uint8_t* d_rgb, int rows, int cols,
                                const int32_t* depth_idx, const int32_t* rgb_idx /* NULL = depth_idx */, int n,
                                const float* Tcw, float fx, float fy, float cx, float cy);
/* Host-buffer form for a batch of keyframes as the reference's callers hold them: the sensor's CV_16U depth images
 * (imDepth before convertTo, src/Tracking.cc:353-367; depth_factor = 1.0f / DepthMapFactor) and the colour images
 * handed to insertKeyFrame (src/Tracking.cc:1889), depth_u16[n][rows][cols], rgb[n][rows][cols][3], Tcw[n][16].
 * Uploads, the conversion and the n inserts are enqueued on the map's stream and the call returns; the host buffers
 * (page-locked for a truly asynchronous upload) must stay untouched until ocm_sync(h). */
int ocm_insert_keyframes_u16(ocm_t* h, const uint16_t* depth_u16, const uint8_t* rgb, int rows, int cols, int n,
                             float depth_factor, const float* Tcw, float fx, float fy, float cx, float cy);
/* World-frame points of the LAST inserted keyframe after gating + leaf filter + transform (xyz f32 x n,
 * unordered: compare as sets). */
int ocm_last_points(ocm_t* h, float* xyz, uint8_t* rgb, int cap, int* n);
/* Leaf export: keys (3 x u16 per leaf, octomap OcTreeKey), log-odds f32; unordered. */
int64_t ocm_num_leaves(ocm_t* h);
int ocm_export_leaves(ocm_t* h, uint16_t* keys, float* logodds, uint8_t* rgb, int64_t cap, int64_t* n);
int ocm_query(ocm_t* h, const float xyz[3], float* logodds, int* found);
/* Multi-GPU map merge: per-voxel clamp-add summaries f(x)=min(max(x+a,L),H) of everything
 * inserted since the last reset; device buffers, n entries: key u64 (x | y<<16 | z<<32), a/L/H f32. */
int64_t ocm_summary_count(ocm_t* h);
int ocm_export_summaries_device(ocm_t* h, uint64_t* d_keys, float* d_a, float* d_lo, float* d_hi, int64_t cap,
                                int64_t* n);
/* Compose the summaries of a LATER shard onto this map (apply in keyframe order: shard 0, 1, ...). */
int ocm_apply_summaries_device(ocm_t* h, const uint64_t* d_keys, const float* d_a, const float* d_lo,
                               const float* d_hi, int64_t n);
/* The same with a per-pixel ground label batch d_label[F][rows][cols] (NULL = all non-ground); keyframe i reads label
 * image label_idx[i] (NULL = depth_idx).  Labels stand in for the reference's RANSAC floor split
 * (perfect/src/MapDrawer.cc:686-774): ground points only clear space along their rays (:961-969). */
int ocm_insert_keyframes_labeled_device(ocm_t* h, const float* d_depth, const uint8_t* d_rgb, const uint8_t* d_label,
                                        int rows, int cols, const int32_t* depth_idx, const int32_t* rgb_idx,
                                        const int32_t* label_idx, int n, const float* Tcw, float fx, float fy, float cx,
                                        float cy);
int ocm_insert_keyframes_u16_labeled(ocm_t* h, const uint16_t* depth_u16, const uint8_t* rgb, const uint8_t* label, int rows,
                                     int cols, int n, float depth_factor, const float* Tcw, float fx, float fy, float cx,
                                     float cy);
/* ocm_merge_nccl (SURVEY §8(b),(e)): the one exchange step of the frame-sharded mode.  Every rank has inserted its own
 * keyframes (a contiguous range, rank order = keyframe order) since the last merge; the call all-gathers the per-voxel
 * clamp-add summaries of that epoch (24-byte records: key u64, a, lo, hi f32, colour u32) over NCCL / NVLink and replays
 * the shards of ALL ranks in rank order on the values of the last merge, so every rank ends with the map a single process
 * inserting all keyframes in order would hold (log-odds equal up to float re-association, << 1e-5).  One host
 * synchronisation (the counts); everything else is enqueued on the map's stream.  comm = ncclComm_t (NULL: single rank,
 * the call just closes the epoch); stream = a cudaStream_t the merge is ordered after / before (NULL: the map's own). */
typedef struct {
  int world, rank;
  int64_t records_sent, records_total;     /* voxels this rank sent / all ranks together */
  int64_t bytes_sent, bytes_received;      /* 24 B per record */
} OcmMergeStats;
int ocm_merge_nccl(ocm_t* h, void* nccl_comm, void* stream, OcmMergeStats* stats);
/* Helpers to obtain an ncclComm_t without linking NCCL: rank 0 calls ocm_nccl_unique_id and hands the 128 bytes to the
 * other ranks through whatever plumbing the host has (torch.distributed.broadcast in the Python mirror). */
int ocm_nccl_unique_id(uint8_t id[128]);
int ocm_nccl_comm_create(const uint8_t id[128], int rank, int world, int device, void** nccl_comm);
int ocm_nccl_comm_destroy(void* nccl_comm);
/* Counters for bench.py's B_map: points = P of the LAST round of keyframes (at most 32) a batch insert ran (points that
 * survived gates + leaf filter, summed over the round's keyframes); voxel_updates = CUMULATIVE number of (voxel, keyframe)
 * log-odds updates since the map was created (U summed over every keyframe so far). */
int ocm_last_batch_stats(ocm_t* h, int64_t* points, int64_t* voxel_updates);
int ocm_sync(ocm_t* h);
void* ocm_stream(ocm_t* h);
long long ocm_launch_count(const ocm_t* h);

/* ---------------------------------------------------------------------------------------------------------------------
 * T-variant dense map: the accumulated colour cloud of PointCloudMapping::viewer (src/pointcloudmapping.cc:395-500).
 *   gcm_add_keyframe[_device]  generatePointCloud (:131-194: every pixel, no depth gate, world frame through Tcw^-1 in
 *                              double) + removeNaNFromPointCloud + "*globalMap += *out_pt" (:482-485); depth f32 metres,
 *                              colour image BGR u8 [rows][cols][3]
 *   gcm_refilter               "voxel.setInputCloud(globalMap); voxel.filter(*tmp); globalMap->swap(*tmp)" (:490-493),
 *                              VoxelGrid leaf = the constructor's resolution (:40)
 *   gcm_export                 the points of globalMap (what viewer / global_color.pcd receive, :496-521), cell order
 * B200ORB_EGEOM from gcm_refilter: the cell index space overflows int for this leaf (PCL leaves the cloud unfiltered). */
typedef struct gcm gcm_t;
int gcm_create(float leaf, int device, gcm_t** out);
void gcm_destroy(gcm_t* h);
int gcm_add_keyframe(gcm_t* h, const float* depth, const uint8_t* bgr, int rows, int cols, const float Tcw[16], float fx,
                     float fy, float cx, float cy);
int gcm_add_keyframe_device(gcm_t* h, const float* d_depth, const uint8_t* d_bgr, int rows, int cols, const float Tcw[16],
                            float fx, float fy, float cx, float cy);
int gcm_refilter(gcm_t* h);
long long gcm_size(const gcm_t* h);
int gcm_export(gcm_t* h, float* xyz, uint8_t* rgb, long long cap, long long* n);
int gcm_sync(gcm_t* h);
long long gcm_launch_count(const gcm_t* h);

#ifdef __cplusplus
}
#endif
#endif /* B200ORB_H_ */
