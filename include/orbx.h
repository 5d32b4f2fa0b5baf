/* orbx.h — C ABI of liborbx.so: B200-native (sm_100a) ORB extraction and Hamming matching.
 *
 * This is the drop-in boundary for the hot path of OpenMAVIS (an ORB-SLAM3 fork).  The reference has
 * no FFI layer: the path is two C++ classes, ORB_SLAM3::ORBextractor (include/ORBextractor.h:28-84,
 * src/ORBextractor.cc) and ORB_SLAM3::ORBmatcher (include/ORBmatcher.h:17-123, src/ORBmatcher.cc), plus
 * the cv::BFMatcher call in Frame (src/Frame.cc:24,1427,1483).  The header-compatible shims in
 * openmavis_b200/csrc/host/ are written on top of exactly these entry points.
 *
 * Conventions: plain C types only; opaque handles; caller-owned output buffers with an explicit capacity;
 * every function returns ORBX_OK (0) or a negative orbx_status and never throws; orbx_last_error() gives
 * the message of the calling thread's last failure.  A handle may be used by one thread at a time;
 * different handles are independent (one CUDA stream set per handle).  There is NO CPU fallback: without
 * a CUDA device every compute entry point fails with ORBX_ERR_NO_DEVICE.
 *
 * All file:line cites are relative to the reference tree.
 */
#ifndef ORBX_H
#define ORBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum orbx_status {
    ORBX_OK = 0,
    ORBX_ERR_INVALID = -1,     /* bad argument */
    ORBX_ERR_NO_DEVICE = -2,   /* no CUDA device / device index out of range */
    ORBX_ERR_CUDA = -3,        /* CUDA runtime error (message in orbx_last_error) */
    ORBX_ERR_CAPACITY = -4,    /* caller buffer too small / batch larger than the handle was created for */
    ORBX_ERR_EMPTY = -5,       /* empty image: the reference returns -1 (src/ORBextractor.cc:994-995) */
    ORBX_ERR_ARGUMENT = -6     /* orbx_extract only: bad argument (its -1 is the reference's return value for an empty image) */
} orbx_status;

/* Field-for-field cv::KeyPoint (pt.x, pt.y, size, angle, response, octave, class_id): 28 bytes. */
typedef struct orbx_kp {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbx_kp;

/* Per-image extractor settings: the constructor arguments of ORBextractor that may differ between the
 * cameras of a rig (src/Tracking.cc:572-587: side cameras get 2*nFeatures) and the vLappingArea argument
 * of operator() (src/ORBextractor.cc:987-992). */
typedef struct orbx_image_params {
    int32_t nfeatures;
    int32_t ini_th_fast;
    int32_t min_th_fast;
    int32_t lap0, lap1;        /* vLappingArea[0], vLappingArea[1] */
} orbx_image_params;

const char *orbx_last_error(void);
int orbx_device_count(void);
/* Library / build identification, e.g. "orbx 0.1 sm_100a". */
const char *orbx_version(void);

/* ---------------------------------------------------------------------------------------------------
 * Extraction — replaces ORBextractor::ORBextractor (src/ORBextractor.cc:351-414) and
 * ORBextractor::operator() (src/ORBextractor.cc:987-1071) with everything under it: ComputePyramid
 * (:1073-1104), ComputeKeyPointsOctTree (:704-823), DistributeOctTree (:496-702), IC_Angle (:19-43),
 * GaussianBlur + computeOrbDescriptor (:1035-1036, :46-90) and the lapping-area partition (:1024-1068).
 *
 * A handle is created for one image geometry (width x height), pyramid shape (nlevels, scale_factor) and
 * a maximum batch size; one call processes up to max_images images (all cameras of a rig frame, or many
 * rig frames) in one launch set.  fma_mode: 0 = every float op of the descriptor rotation rounded
 * separately; 1 = contracted the way g++ -O3 -march=native contracts src/ORBextractor.cc:56-57.
 * ------------------------------------------------------------------------------------------------- */
typedef struct orbx_extractor orbx_extractor;

int orbx_extractor_create(orbx_extractor **out, int device, int width, int height, int max_images,
                          int max_nfeatures, float scale_factor, int nlevels, int fma_mode);
void orbx_extractor_destroy(orbx_extractor *ex);

/* Launch all work of this handle on an existing CUDA stream (cudaStream_t passed as void*); NULL restores
 * the handle's own stream. */
int orbx_extractor_set_stream(orbx_extractor *ex, void *cuda_stream);

/* Rows of the per-image output buffers the caller must provide (>= nfeatures plus the quadtree overshoot). */
int orbx_extractor_capacity(const orbx_extractor *ex);

/* Per-level tables = the getters of ORBextractor.h:40-50 (GetScaleFactors, GetInverseScaleFactors,
 * GetScaleSigmaSquares, GetInverseScaleSigmaSquares) and the level sizes of ComputePyramid. Any pointer may be NULL. */
int orbx_extractor_tables(const orbx_extractor *ex, float *scale, float *inv_scale, float *sigma2,
                          float *inv_sigma2, int *level_w, int *level_h);
/* mnFeaturesPerLevel for a given nfeatures (src/ORBextractor.cc:378-391). */
int orbx_features_per_level(int nfeatures, float scale_factor, int nlevels, int *out);

/* Host-buffer call (what ORBextractor::operator() binds to).  images[i] points to a CV_8UC1 image of the
 * handle's geometry with row stride strides[i] bytes in HOST memory.  Outputs, all HOST memory:
 *   kps  [n_images][cap]      keypoints, mono block first then the reversed lapping block
 *   desc [n_images][cap][32]  descriptor rows in the same order
 *   n_kp [n_images]           number of keypoints     n_mono [n_images]  return value of operator()
 * The call uploads, computes, downloads and returns when the outputs are valid. */
int orbx_extract_batch(orbx_extractor *ex, int n_images, const uint8_t *const *images, const int *strides,
                       const orbx_image_params *params, orbx_kp *kps, uint8_t *desc, int cap, int *n_kp,
                       int *n_mono);

/* Asynchronous form of orbx_extract_batch for throughput pipelines: submit enqueues the uploads, the kernels and the
 * downloads of one batch and returns; orbx_extract_wait blocks until the outputs of the last submit are valid.  The host
 * buffers (images, outputs) must stay valid until the wait and should be pinned, or the copies are not asynchronous.  Two handles
 * used alternately (submit A, submit B, wait A, submit A, ...) overlap the copies of one batch with the kernels of the other. */
int orbx_extract_batch_submit(orbx_extractor *ex, int n_images, const uint8_t *const *images, const int *strides,
                              const orbx_image_params *params, orbx_kp *kps, uint8_t *desc, int cap, int *n_kp,
                              int *n_mono);
int orbx_extract_wait(orbx_extractor *ex);

/* Single image, reference-shaped: returns n_mono (>= 0) like operator(), or -1 for an empty image,
 * or another negative orbx_status. */
int orbx_extract(orbx_extractor *ex, const uint8_t *image, int stride, const orbx_image_params *params,
                 orbx_kp *kps, uint8_t *desc, int cap, int *n_kp);

/* Device-resident call: d_images is DEVICE memory holding n_images tightly packed width x height images.
 * Enqueues the whole pipeline on the handle's stream and returns without synchronising; results stay in
 * the handle's device buffers until fetched. */
int orbx_extract_batch_device(orbx_extractor *ex, int n_images, const uint8_t *d_images,
                              const orbx_image_params *params);
/* Copies the results of the last (device or host) call to host buffers; synchronises the stream. */
int orbx_extract_fetch(orbx_extractor *ex, orbx_kp *kps, uint8_t *desc, int cap, int *n_kp, int *n_mono);
/* Device pointers of the last results (for device-side consumers such as the matchers):
 * kps [max_images][cap], desc [max_images][cap][32], counts: n_kp[max_images], n_mono[max_images]. */
int orbx_extract_device_results(orbx_extractor *ex, const orbx_kp **d_kps, const uint8_t **d_desc,
                                const int **d_n_kp, const int **d_n_mono);

/* mvImagePyramid[level] of image `image` from the last call (public member read by
 * src/Frame.cc:1063,1147,1163).  with_border=0: the w x h level; 1: the (w+38) x (h+38) buffer with the
 * 19-px BORDER_REFLECT_101 frame.  Downloaded lazily, on request. */
int orbx_extractor_pyramid_level(orbx_extractor *ex, int image, int level, int with_border, uint8_t *dst,
                                 int dst_stride);

/* Frame::ComputeStereoMatches (src/Frame.cc:1057-1219) for a rectified pair whose two images were extracted by the LAST call
 * of this handle (plane indices img_left / img_right): row-band Hamming matching (TH_HIGH, octave +-1, disparity range from
 * mb / mbf), 11x11 SAD refinement on the device-resident pyramid levels, parabola sub-pixel fit, depth, median-based outlier
 * removal.  u_right[i] / depth[i] (i < n_kp[img_left], HOST, capacity cap >= orbx_extractor_capacity()) = mvuRight / mvDepth,
 * -1 where there is no match; *n_matches = number of keypoints that keep a depth.  mb = baseline, mbf = baseline * fx. */
int orbx_compute_stereo_matches(orbx_extractor *ex, int img_left, int img_right, float mb, float mbf, float *u_right,
                                float *depth, int cap, int *n_matches);
/* Same with the two images held by two handles of identical geometry (mpORBextractorLeft / mpORBextractorRight). */
int orbx_compute_stereo_matches2(orbx_extractor *ex_left, int img_left, orbx_extractor *ex_right, int img_right, float mb, float mbf,
                                 float *u_right, float *depth, int cap, int *n_matches);

/* Stage intermediates of the last call, for stage-by-stage parity tests:
 * blurred level (w x h), FAST candidates in vToDistributeKeys order (border-relative coordinates,
 * src/ORBextractor.cc:784-791), per-level keypoints after DistributeOctTree + IC_Angle (level coordinates). */
int orbx_debug_blur_level(orbx_extractor *ex, int image, int level, uint8_t *dst, int dst_stride);
int orbx_debug_candidates(orbx_extractor *ex, int image, int level, int *xs, int *ys, int *scores, int cap,
                          int *n);
int orbx_debug_level_keypoints(orbx_extractor *ex, int image, int level, orbx_kp *kps, int cap, int *n);

/* Per-stage device time of the last call, measured with CUDA events on the handle's stream when
 * profiling is enabled (adds event records between stages; off by default).
 * ms[0..ORBX_NUM_STAGES): pyramid, blur, fast, octree, slots, orient+descriptor, total. */
#define ORBX_NUM_STAGES 7
int orbx_extractor_set_profiling(orbx_extractor *ex, int enabled);
int orbx_extractor_stage_ms(orbx_extractor *ex, float *ms);
/* FAST counters accumulated over the profiled calls since profiling was switched on (profiling mode only):
 * out6 = tested pixels, quick-test survivors, corners (score >= threshold), kept after NMS, retried cells, tiles. */
int orbx_extractor_fast_stats(orbx_extractor *ex, unsigned long long *out6);
/* Number of kernel launches enqueued by the last call. */
int orbx_extractor_last_launches(const orbx_extractor *ex);

/* ---------------------------------------------------------------------------------------------------
 * Matching.
 * ------------------------------------------------------------------------------------------------- */

/* ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2577-2591) / FORB::distance
 * (Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101) for n independent pairs a[i], b[i] (32-byte rows, HOST). */
int orbx_descriptor_distance(int device, const uint8_t *a, const uint8_t *b, int n, int *dist);

typedef struct orbx_matcher orbx_matcher;
int orbx_matcher_create(orbx_matcher **out, int device);
void orbx_matcher_destroy(orbx_matcher *m);
int orbx_matcher_set_stream(orbx_matcher *m, void *cuda_stream);

/* cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, matches, 2) (src/Frame.cc:1427, 1483): for each of the
 * nq query rows the two nearest of the nt train rows, ordered (distance asc, trainIdx asc).
 * idx/dist are [nq][2]; an entry that does not exist (nt < 2) is -1.  HOST pointers. */
int orbx_knn2(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int nt, int *idx, int *dist);
/* Same with DEVICE pointers, asynchronous on the matcher's stream (no host synchronisation). */
int orbx_knn2_device(orbx_matcher *m, const uint8_t *d_query, int nq, const uint8_t *d_train, int nt, int *d_idx,
                     int *d_dist);
/* Sharded brute-force 2-NN: the train set is the concatenation of nparts parts that live in DEVICE memory of
 * this or of PEER GPUs (pointers obtained with orbx_shared_open): the matcher reads the peers' rows in place over NVLink —
 * there is no all-gather step.  Indices are global (offset of part p = sum of part_rows[0..p)).  Asynchronous on the
 * matcher's stream; the caller orders it after the peers have written their rows (a barrier). */
int orbx_knn2_device_parts(orbx_matcher *m, const uint8_t *d_query, int nq, const uint8_t *const *d_train_parts,
                           const int *part_rows, int nparts, int *d_idx, int *d_dist);
/* The same with the barrier and the overlap built in.  d_flags[p] = device pointer (peer-mapped for p != self) to part p's 32-bit
 * epoch word: the owner of part p calls orbx_shared_publish(flag, k) on its stream after its rows of step k are in place, and this
 * call (with epoch = k) waits for the peers' flags ON THE DEVICE, inside the kernel that pulls their rows over NVLink on an auxiliary
 * stream, while the main stream already matches the own part (index `self`, or -1) in place: no collective, no host synchronisation.
 * With two shared buffers per rank (step k uses buffer k & 1) no further barrier is needed: a rank can only see epoch k of a peer
 * after that peer has finished pulling step k - 1.  A peer that never publishes makes the pull kernel trap after timeout_s (<= 0: 10 s)
 * — the next CUDA call reports the error; nothing spins forever. */
int orbx_knn2_device_parts_sync(orbx_matcher *m, const uint8_t *d_query, int nq, const uint8_t *const *d_train_parts,
                                const int *part_rows, int nparts, int self, const uint32_t *const *d_flags, uint32_t epoch,
                                double timeout_s, int *d_idx, int *d_dist);
/* Stores `epoch` into a flag word with release semantics at system scope, ordered after everything queued on cuda_stream. */
int orbx_shared_publish(int device, void *d_flag, uint32_t epoch, void *cuda_stream);
/* Device buffers that other processes on the same node can map (cudaIpc): alloc in the owner (handle = 64 opaque bytes to
 * send to the peers), open / close in a peer, free in the owner.  A new buffer is zero-filled. */
int orbx_shared_alloc(int device, size_t bytes, void **d_ptr, void *ipc_handle64);
int orbx_shared_open(int device, const void *ipc_handle64, void **d_ptr);
int orbx_shared_close(int device, void *d_ptr);
int orbx_shared_free(int device, void *d_ptr);
/* device-to-device copy on a CUDA stream (fills a shared buffer from any device memory). */
int orbx_device_copy(int device, void *d_dst, const void *d_src, size_t bytes, void *cuda_stream);

/* Lowe ratio test exactly as the callers evaluate it (src/Frame.cc:1434 ratio 0.7; :1491 ratio 0.8):
 * keep[i] = two neighbours exist && (float)d0 < (float)d1 * ratio (product in double). HOST, pure host code. */
int orbx_ratio_test(const int *dist, int nq, double ratio, uint8_t *keep);

/* Candidate-list matching: the Hamming core shared by the guided matchers (SearchByProjection overloads,
 * Fuse, SearchBySim3, SearchForInitialization; src/ORBmatcher.cc:72-114 and its clones).  Query q is
 * compared with train rows cand[row_ptr[q] .. row_ptr[q+1]) in that order.  Per query:
 *   best_idx / best_dist    first minimum in candidate order (strict <), -1 / 256 if the list is empty
 *   second_dist             ORB-SLAM second-best rule: `else if (dist < bestDist2)` (src/ORBmatcher.cc:103)
 *   best_pos / second_pos   positions inside the candidate list (for level bookkeeping), -1 if none
 * HOST pointers. */
int orbx_match_csr(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int nt, const int *row_ptr,
                   const int *cand, int *best_idx, int *best_dist, int *second_dist, int *best_pos,
                   int *second_pos);
/* All candidate distances, CSR-aligned: dist[k] = hamming(query[q], train[cand[k]]) for k in row q. */
int orbx_match_csr_dists(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int nt,
                         const int *row_ptr, const int *cand, int *dist);

/* ORBmatcher::SearchForTriangulation Hamming core (src/ORBmatcher.cc:1218-1427, bCoarse semantics): the two
 * vocabulary FeatureVectors as CSR (ascending node ids, index lists); skip1/skip2 mark keypoints that
 * already have a MapPoint (or fail bOnlyStereo).  match12[n1] = idx2 or -1: minimum distance <= TH_LOW(50),
 * LAST minimum on ties (src/ORBmatcher.cc:1274).  Returns the number of matches in *nmatches. HOST pointers. */
int orbx_search_for_triangulation(orbx_matcher *m, int nn1, const int *nodes1, const int *ptr1, const int *idx1,
                                  int nn2, const int *nodes2, const int *ptr2, const int *idx2,
                                  const uint8_t *desc1, int n1, const uint8_t *desc2, int n2, const uint8_t *skip1,
                                  const uint8_t *skip2, int *match12, int *nmatches);

/* ORBmatcher::SearchByBoW(KeyFrame *pKF, Frame &F, vpMapPointMatches) (src/ORBmatcher.cc:349-666), the whole function: the two
 * FeatureVectors as CSR (ascending node ids, feature index lists in insertion order); skip_kf[i] = the KeyFrame feature has no
 * usable MapPoint (:377-380); angle_kf / angle_f = keypoint angles by feature index (only read with check_orientation);
 * ncam_f = 1 for F.Nleft == -1, else 2 or 4 with the camera blocks cam_off_f[0..ncam_f] of F's feature indices (:406-452).
 * Per shared node and KeyFrame feature, in order: best / second-best distance per camera over the F features not matched yet;
 * camera 0 needs dist <= TH_LOW and the mfNNratio test, cameras 1-3 only dist <= TH_LOW (the reference's `|| true`), all nested in
 * camera 0's TH_LOW test (:456); then the rotation-histogram filter (:637-663).  match_f[n_f] = KeyFrame feature index whose
 * MapPoint goes to vpMapPointMatches[i] (or -1), *nmatches = the return value.  HOST pointers. */
int orbx_search_by_bow_frame(orbx_matcher *m, int nn_kf, const int *nodes_kf, const int *ptr_kf, const int *idx_kf, int nn_f,
                             const int *nodes_f, const int *ptr_f, const int *idx_f, const uint8_t *desc_kf, int n_kf,
                             const uint8_t *desc_f, int n_f, const uint8_t *skip_kf, const float *angle_kf, const float *angle_f,
                             int ncam_f, const int *cam_off_f, float nnratio, int check_orientation, int *match_f, int *nmatches);
/* ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vpMatches12) (src/ORBmatcher.cc:1006-1129): skip1 / skip2 = feature
 * without usable MapPoint or beyond N (:1036-1044, 1052-1061); best < TH_LOW (strict) and the mfNNratio test; vbMatched2 keeps a
 * second-list feature from being matched twice; rotation-histogram filter.  match12[n1] = idx2 (or -1).  HOST pointers. */
int orbx_search_by_bow_keyframes(orbx_matcher *m, int nn1, const int *nodes1, const int *ptr1, const int *idx1, int nn2, const int *nodes2,
                                 const int *ptr2, const int *idx2, const uint8_t *desc1, int n1, const uint8_t *desc2, int n2,
                                 const uint8_t *skip1, const uint8_t *skip2, const float *angle1, const float *angle2, float nnratio,
                                 int check_orientation, int *match12, int *nmatches);

/* Frame::AssignFeaturesToGrid + Frame::GetFeaturesInArea (src/Frame.cc:541-582, 890-967, 969-978) for one
 * camera: builds the 64x48 grid of `kps` on the device and answers nq window queries, producing CSR
 * candidate lists in the reference's order (cell column-major, insertion order inside a cell).
 * row_ptr[nq+1]; cand has capacity cand_cap; *n_cand receives the total (may exceed cand_cap -> ORBX_ERR_CAPACITY). */
int orbx_grid_candidates(orbx_matcher *m, const orbx_kp *kps, int n, float min_x, float max_x, float min_y,
                         float max_y, const float *qx, const float *qy, const float *qr, const int *min_level,
                         const int *max_level, int nq, int *row_ptr, int *cand, int cand_cap, int *n_cand);

/* ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints)
 * (src/ORBmatcher.cc:23-340) over plain arrays; argument meaning as in oracle/orb_oracle.cpp
 * (orc_search_by_projection).  Candidate generation, Hamming distances, best/second selection and the ordered
 * claim resolution (a keypoint claimed by an earlier map point is skipped by later ones, :77-79) all run on
 * the device; assigned[N] receives the map-point index written to F.mvpMapPoints (or -1).
 * u_right[N] = F.mvuRight and proj_xr[nmp] = MapPoint::mTrackProjXR: the rectified-stereo gate of :81-85 (only applied with
 * ncam == 1, i.e. F.Nleft == -1; both NULL = monocular frame without right observations).
 * Returns ORBX_ERR_INVALID for NULL inputs or a predicted level outside [0, nlevels) on an in-view entry.  HOST pointers. */
int orbx_search_by_projection(orbx_matcher *m, int ncam, const int *cam_off, const orbx_kp *kps, const uint8_t *desc,
                              const float *grid_bounds /* [ncam][4] minX,maxX,minY,maxY */, const float *scale_factors,
                              int nlevels, const uint8_t *occupied, const int *l2r, const int *r2l, const float *u_right, int nmp,
                              const uint8_t *in_view, const float *proj_x, const float *proj_y, const int *level,
                              const float *view_cos, const float *proj_xr, const float *depth, const uint8_t *bad, const uint8_t *has_obs,
                              const uint8_t *mp_desc, float th, int far_points, float th_far_points, float nnratio,
                              int *assigned, int *nmatches);

/* ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (src/ORBmatcher.cc:1985-2413): the
 * frame-to-frame tracking matcher.  Frame side as in orbx_search_by_projection; u_right = CurrentFrame.mvuRight (only read when
 * ncam == 1; may be NULL).  Per last-frame keypoint i with a usable map point (valid[i]: pMP set, not an outlier, invzc >= 0 and
 * the camera-0 projection inside the image bounds, :2008-2026): proj_x/proj_y[i][c] = the projection searched in camera c (the
 * caller passes what the reference passes — for the side cameras that is the LEFT projection, :2250-2271), proj_ur[i] = uv.x -
 * mbf * invzc, last_octave / last_angle = octave and angle of the last frame's keypoint, has_obs[i] = Observations() > 0,
 * pt_desc = GetDescriptor().  level_gate = ORBX_GATE_FORWARD / _BACKWARD for bForward / bBackward (:2003-2004), else _AROUND.
 * Window th * sf[octave], level gate by direction, best Hamming distance <= max_dist (TH_HIGH = 100 here) among the keypoints not claimed by an earlier point; an empty camera-0 window skips the
 * point's other cameras (:2062); with check_orientation the rotation-histogram filter (:2392-2410) clears the assignments outside
 * the three fullest bins.  assigned[N] = i written to CurrentFrame.mvpMapPoints (or -1), *nmatches = the return value.
 * The relocalisation overload SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, sAlreadyFound, th, ORBdist) (:2415-2535) is the
 * same call with ncam = 1 (it searches camera 0 only), valid[i] = its own early `continue`s (:2433-2459), last_octave =
 * nPredictedLevel, last_angle = pKF->mvKeysUn[i].angle, occupied = (mvpMapPoints[idx] != NULL), has_obs = NULL (every claim is
 * visible, :2478), ORBX_GATE_AROUND and max_dist = ORBdist.  The loop-closing overloads SearchByProjection(KeyFrame *pKF, Sim3 Scw,
 * vpPoints, vpMatched, th, ratioHamming, cameraID) (:668-774) and its vpMatchedKF twin (:776-893) are the same call on the chosen
 * camera's block of the KeyFrame (ncam = 1): occupied = (vpMatched[idx] != NULL), ORBX_GATE_BELOW, check_orientation = 0,
 * max_dist = (int)floorf(TH_LOW * ratioHamming) (:767).  HOST pointers. */
#define ORBX_GATE_AROUND 0   /* [l - 1, l + 1] */
#define ORBX_GATE_FORWARD 1  /* [l, inf)       */
#define ORBX_GATE_BACKWARD 2 /* [0, l]         */
#define ORBX_GATE_BELOW 3    /* [l - 1, l]     */
int orbx_search_by_projection_last(orbx_matcher *m, int ncam, const int *cam_off, const orbx_kp *kps, const uint8_t *desc,
                                   const float *grid_bounds, const float *scale_factors, int nlevels, const uint8_t *occupied,
                                   const float *u_right, int npts, const uint8_t *valid, const float *proj_x, const float *proj_y,
                                   const float *proj_ur, const int *last_octave, const float *last_angle, const uint8_t *has_obs,
                                   const uint8_t *pt_desc, float th, int level_gate, int check_orientation, int max_dist,
                                   int *assigned, int *nmatches);
/* After orbx_search_by_projection_last: cleared[idx] = 1 for the keypoints that were claimed and then cleared by the rotation filter —
 * the reference writes NULL into CurrentFrame.mvpMapPoints there (src/ORBmatcher.cc:2404), while a keypoint no point claimed keeps what
 * it held; assigned[] reports -1 for both.  n = the keypoint count of that call.  HOST pointer. */
int orbx_search_by_projection_last_cleared(const orbx_matcher *m, uint8_t *cleared, int n);

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:895-1004), whole
 * function: kps / desc = mvKeysUn / mDescriptors of the two frames, grid_bounds2 = F2's {mnMinX, mnMaxX, mnMinY, mnMaxY};
 * prev_matched [n1][2] = vbPrevMatched (in: search centres; out: updated with the matched F2 positions, :992-995).  Level-0
 * keypoints of F1 only, window = windowSize at level 0, per F2 keypoint the smallest distance so far wins and evicts the previous
 * owner (vMatchedDistance / vnMatches21, :927, 941-947), TH_LOW, bestDist < bestDist2 * mfNNratio, rotation-histogram filter.
 * matches12[n1] = vnMatches12, *nmatches = the return value.  HOST pointers. */
int orbx_search_for_initialization(orbx_matcher *m, const orbx_kp *kps1, const uint8_t *desc1, int n1, const orbx_kp *kps2,
                                   const uint8_t *desc2, int n2, const float *grid_bounds2, float *prev_matched, int window_size,
                                   float nnratio, int check_orientation, int *matches12, int *nmatches);

/* Window search without the "already matched" rule — the search loop of Fuse (src/ORBmatcher.cc:1458-1640 and :1642-1769) and of
 * both directions of SearchBySim3 (:1771-1983) over one camera of a KeyFrame: kps / desc = that camera's keypoints (camera-local
 * indices), grid_bounds = {mnMinX, mnMaxX, mnMinY, mnMaxY}.  Per query q with valid[q] (the caller's projection tests): the
 * keypoints of KeyFrame::GetFeaturesInArea(qx, qy, radius) (src/KeyFrame.cc:771-834) with skip[idx] == 0, octave in
 * [min_level, max_level], and — when inv_level_sigma2 is given (Fuse, :1571-1592) — squared reprojection error * invSigma2[octave]
 * <= 5.99, or <= 7.8 including (q_ur - u_right[idx])^2 where u_right[idx] >= 0.  best_idx[q] = first minimum of the Hamming
 * distance (-1 if none), best_dist[q] (256 if none); the caller applies its threshold (TH_LOW / TH_HIGH) and graph updates.
 * skip, u_right, q_ur, inv_level_sigma2 may be NULL.  HOST pointers. */
int orbx_window_best(orbx_matcher *m, const orbx_kp *kps, const uint8_t *desc, int n, const float *grid_bounds, const uint8_t *skip,
                     const float *u_right, const float *inv_level_sigma2, int nlevels, int nq, const uint8_t *valid, const float *qx,
                     const float *qy, const float *q_ur, const float *radius, const int *min_level, const int *max_level,
                     const uint8_t *q_desc, int *best_idx, int *best_dist);
/* ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:1771-1983): both window searches (orbx_window_best semantics, level gate
 * [l - 1, l], best <= TH_HIGH) and the agreement check (:1968-1980).  Side 1 queries = the map points of KeyFrame 1 projected into
 * KeyFrame 2 (valid1[i1] = 0 for !pMP, vbAlreadyMatched1, bad, behind the camera, outside the image or the scale range), searched in
 * KeyFrame 2's keypoints with mp_desc1[i1] = pMP->GetDescriptor(); side 2 likewise.  match12[n1] = idx2 with vnMatch1[i1] == idx2 && vnMatch2[idx2] == i1, else -1;
 * *nfound = the return value.  HOST pointers. */
int orbx_search_by_sim3(orbx_matcher *m, const orbx_kp *kps1, const uint8_t *desc1, int n1, const float *grid_bounds1,
                        const orbx_kp *kps2, const uint8_t *desc2, int n2, const float *grid_bounds2, const uint8_t *valid1,
                        const float *u1, const float *v1, const float *radius1, const int *level1, const uint8_t *mp_desc1,
                        const uint8_t *valid2, const float *u2, const float *v2, const float *radius2, const int *level2,
                        const uint8_t *mp_desc2, int *match12, int *nfound);

/* Fisheye stereo post-filter.  Camera pair of a Kannala-Brandt rig: cam = {fx, fy, cx, cy, k0, k1, k2, k3}
 * (KannalaBrandt8::mvParameters), R12 / t12 = Frame::mRlr / mtlr (row-major rotation, translation), precision = the Newton
 * stopping threshold of KannalaBrandt8::unproject (1e-6 in the reference). */
typedef struct orbx_kb8_stereo {
    float cam1[8], cam2[8];
    float R12[9], t12[3];
    float precision1, precision2;
} orbx_kb8_stereo;

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (src/ORBmatcher.cc:1131-1456), the WHOLE
 * function incl. bCoarse = false (what LocalMapping::CreateNewMapPoints calls, src/LocalMapping.cc:465-468): descriptor gate, the
 * epipole-distance reject of the monocular case (:1292-1299), the camera-pair choice whose (pCamera1, pCamera2, R12, t12) is carried
 * across candidates when a pair of cameras has no branch (:1301-1378), pCamera1->epipolarConstrain (:1380-1387), the last-minimum rule,
 * the rotation histogram (only with check_orientation).
 *   ncam            1 (NLeft == -1: monocular / rectified stereo, no mpCamera2), 2 (left + right) or 4 (+ side-left, side-right);
 *                   camera c of keyframe k owns the feature indices [cam_offk[c], cam_offk[c + 1])
 *   kpsK / descK    keypoints (pt, octave, angle are read) and descriptors by feature index; has_mpK[idx] = GetMapPoint(idx) != NULL
 *   u_rightK        mvuRight (only read when ncam == 1; NULL = no stereo observations)
 *   FeatureVectors  as CSR: ascending node ids, feature index lists in insertion order
 *   ep              the epipole pKF2->mpCamera->project(T2w * Cw1) (:1141-1147); sf2 = pKF2->mvScaleFactors;
 *                   sigma2_1 / sigma2_2 = mvLevelSigma2 of the two keyframes; nlevels <= 16
 *   pairs[11]       the camera pairs the reference prepares (:1156-1198): [0] = T12 of the ncam == 1 case, [1] ll, [2] lr, [3] rl, [4] rr,
 *                   [5] l-sl, [6] sl-l, [7] sl-sl, [8] r-sr, [9] sr-r, [10] sr-sr.  Pinhole pairs carry F12 = K1^-T [t12]x R12 K2^-1 as
 *                   Pinhole::epipolarConstrain computes it (src/CameraModels/Pinhole.cpp:112-115; evaluated once per pair by the caller
 *                   with the reference's own expression); Kannala-Brandt pairs carry the two parameter vectors, R12, t12 and the
 *                   unproject precision (KannalaBrandt8::epipolarConstrain = TriangulateMatches(...) > 0.0001).  NULL with coarse != 0.
 *                   With ncam >= 2 the reference enters the loop with R12 / t12 uninitialised (:1152-1154): a first candidate whose
 *                   cameras have no branch is tested with pair [1] here.
 * match12[n1] = idx2 or -1 (vMatchedPairs lists them by ascending idx1); *nmatches = the return value.  HOST pointers. */
#define ORBX_CAM_PINHOLE 0
#define ORBX_CAM_KB8 1
typedef struct orbx_epipolar_pair {
    int model;                 /* ORBX_CAM_PINHOLE or ORBX_CAM_KB8 (both cameras of a pair) */
    int reserved;
    float F12[9];              /* pinhole: row-major fundamental matrix */
    struct orbx_kb8_stereo kb8; /* Kannala-Brandt: cam1 / cam2 parameters, R12, t12, unproject precision */
} orbx_epipolar_pair;
int orbx_search_for_triangulation_full(orbx_matcher *m, int ncam, const int *cam_off1, const orbx_kp *kps1, const uint8_t *desc1,
                                       const uint8_t *has_mp1, const float *u_right1, int nn1, const int *nodes1, const int *ptr1,
                                       const int *idx1, const int *cam_off2, const orbx_kp *kps2, const uint8_t *desc2,
                                       const uint8_t *has_mp2, const float *u_right2, int nn2, const int *nodes2, const int *ptr2,
                                       const int *idx2, const float *ep, const float *sf2, const float *sigma2_1, const float *sigma2_2,
                                       int nlevels, int only_stereo, int coarse, int check_orientation, const orbx_epipolar_pair *pairs,
                                       int *match12, int *nmatches);

/* KannalaBrandt8::TriangulateMatches (src/CameraModels/KannalaBrandt8.cpp:319-388) for n independent keypoint pairs:
 * xy1 / xy2 = kp.pt of the two cameras ([n][2]), sigma1 / sigma2 = the two reprojection gates (mvLevelSigma2 of the octaves).
 * depth[i] = z1 > 0, or the reference's reject code (-1 parallax, -2 / -3 behind a camera, -4 / -5 reprojection error);
 * p3d[i] = the triangulated point in camera-1 coordinates (zeros when rejected).  HOST pointers. */
int orbx_kb8_triangulate_matches(orbx_matcher *m, const orbx_kb8_stereo *rig, const float *xy1, const float *xy2, const float *sigma1,
                                 const float *sigma2, int n, float *depth, float *p3d);
/* Frame::ComputeStereoFishEyeMatches (src/Frame.cc:1409-1459, ratio 0.7) / ComputeMultiFishEyeMatches (:1461-1525, ratio 0.8) in
 * one call: brute-force 2-NN of the left lapping block [mono_left, n_left) in the right lapping block [mono_right, n_right),
 * Lowe ratio test, TriangulateMatches, depth > 0.0001.  left_to_right[n_left] / right_to_left[n_right] = mvLeftToRightMatch /
 * mvRightToLeftMatch (-1 = none; a right keypoint chosen by several left ones keeps the last, as the reference's loop does),
 * depth[n_left] = mvDepth (-1), p3d[n_left][3] = mvStereo3Dpoints; counts[0] = nMatches, counts[1] = descMatches (may be NULL).
 * HOST pointers. */
int orbx_stereo_fisheye_matches(orbx_matcher *m, const orbx_kb8_stereo *rig, const orbx_kp *kps_left, const uint8_t *desc_left,
                                int n_left, int mono_left, const orbx_kp *kps_right, const uint8_t *desc_right, int n_right,
                                int mono_right, const float *level_sigma2, int nlevels, double ratio, int *left_to_right,
                                int *right_to_left, float *depth, float *p3d, int *counts);

/* Same with the keypoints and descriptors of the two cameras already in DEVICE memory — the rows the extraction left there
 * (orbx_extract_device_results: d_kps + image * cap, d_desc + image * cap * 32, with n_left / mono_left = the n_kp / n_mono the
 * extraction call returned for the image): no descriptor leaves the GPU between extraction and stereo matching.  The caller orders
 * the call after the extraction (same stream via orbx_matcher_set_stream, or a synchronisation).  Results: HOST pointers. */
int orbx_stereo_fisheye_matches_device(orbx_matcher *m, const orbx_kb8_stereo *rig, const orbx_kp *d_kps_left, const uint8_t *d_desc_left,
                                       int n_left, int mono_left, const orbx_kp *d_kps_right, const uint8_t *d_desc_right, int n_right,
                                       int mono_right, const float *level_sigma2, int nlevels, double ratio, int *left_to_right,
                                       int *right_to_left, float *depth, float *p3d, int *counts);

/* MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:405-483), batched over map points: group g = descriptor rows
 * [ptr[g], ptr[g+1]) of desc (the observations of one map point, in the order the reference collects them).
 * best[g] = index inside the group of the descriptor with the least median Hamming distance to the rest (first minimum,
 * median = sorted row [0.5 * (N - 1)]), -1 for an empty group.  HOST pointers. */
int orbx_distinctive_descriptors(orbx_matcher *m, const uint8_t *desc, const int *ptr, int ngroups, int *best);

/* DBoW2 vocabulary tree (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h): node i has the children child[child_ptr[i] ..
 * child_ptr[i+1]) in the order of Node::children, a 32-byte descriptor, and for leaves word_id / weight; node 0 is the root,
 * levels = m_L.  orbx_vocabulary_transform = the per-feature descent of transform() (:1213-1259) for n descriptors: word_id[i],
 * node_id[i] (node at level m_L - levelsup, as Frame::ComputeBoW uses it with levelsup = 4, src/Frame.cc:980-985) and, if not
 * NULL, weight[i] of the word.  The BowVector / FeatureVector maps are then filled by the caller exactly as :1148-1160 does
 * (addWeight / addFeature in feature order, normalize).  HOST pointers. */
typedef struct orbx_vocabulary orbx_vocabulary;
int orbx_vocabulary_create(orbx_vocabulary **out, int device, int n_nodes, const int *child_ptr, const int *child,
                           const uint8_t *node_desc, const int *word_id, const double *weight, int levels);
void orbx_vocabulary_destroy(orbx_vocabulary *v);
int orbx_vocabulary_transform(orbx_matcher *m, const orbx_vocabulary *v, const uint8_t *desc, int n, int levelsup, int *word_id,
                              int *node_id, double *weight);

/* Launch count / device time of the matcher's last call (CUDA events on its stream). */
/* ---- Guided matchers batched across frames ("report queries/s and latency, batch across frames") -------------------
 * The guided matchers of one frame are latency-bound (a few thousand Hamming distances followed by an ordered, single-CTA claim
 * resolution).  A pool holds `lanes` matcher handles (own stream, events and device buffers each) and as many host worker threads:
 * the calls of a batch are distributed over the lanes and run concurrently, so the device executes the resolvers of up to `lanes` frames
 * side by side while the workers marshal the next ones.  Results are those of the single-frame entry points, call by call.
 * orbx_*_call structs carry exactly the arguments of the corresponding single-frame function; status[i] receives its return value. */
typedef struct orbx_matcher_pool orbx_matcher_pool;
int orbx_matcher_pool_create(orbx_matcher_pool **out, int device, int lanes);
void orbx_matcher_pool_destroy(orbx_matcher_pool *p);
typedef struct orbx_sbp_call {            /* orbx_search_by_projection */
    int ncam; const int *cam_off; const orbx_kp *kps; const uint8_t *desc; const float *grid_bounds; const float *scale_factors; int nlevels;
    const uint8_t *occupied; const int *l2r; const int *r2l; const float *u_right; int nmp; const uint8_t *in_view; const float *proj_x;
    const float *proj_y; const int *level; const float *view_cos; const float *proj_xr; const float *depth; const uint8_t *bad;
    const uint8_t *has_obs; const uint8_t *mp_desc; float th; int far_points; float th_far_points; float nnratio;
    int *assigned; int nmatches;          /* outputs */
} orbx_sbp_call;
typedef struct orbx_sbp_last_call {       /* orbx_search_by_projection_last */
    int ncam; const int *cam_off; const orbx_kp *kps; const uint8_t *desc; const float *grid_bounds; const float *scale_factors; int nlevels;
    const uint8_t *occupied; const float *u_right; int npts; const uint8_t *valid; const float *proj_x; const float *proj_y;
    const float *proj_ur; const int *last_octave; const float *last_angle; const uint8_t *has_obs; const uint8_t *pt_desc; float th;
    int level_gate; int check_orientation; int max_dist;
    int *assigned; int nmatches;          /* outputs */
} orbx_sbp_last_call;
typedef struct orbx_bow_call {            /* orbx_search_by_bow_frame */
    int nn_kf; const int *nodes_kf; const int *ptr_kf; const int *idx_kf; int nn_f; const int *nodes_f; const int *ptr_f; const int *idx_f;
    const uint8_t *desc_kf; int n_kf; const uint8_t *desc_f; int n_f; const uint8_t *skip_kf; const float *angle_kf; const float *angle_f;
    int ncam_f; const int *cam_off_f; float nnratio; int check_orientation;
    int *match_f; int nmatches;           /* outputs */
} orbx_bow_call;
int orbx_search_by_projection_batch(orbx_matcher_pool *p, int ncalls, orbx_sbp_call *calls, int *status);
int orbx_search_by_projection_last_batch(orbx_matcher_pool *p, int ncalls, orbx_sbp_last_call *calls, int *status);
int orbx_search_by_bow_frame_batch(orbx_matcher_pool *p, int ncalls, orbx_bow_call *calls, int *status);

int orbx_matcher_last_ms(orbx_matcher *m, float *ms);
int orbx_matcher_last_launches(const orbx_matcher *m);

/* Peak-rate microbenchmark used as the matching roofline denominator: issues `iters` dependent-free
 * popc.b32 per thread on a full grid and returns giga-popc/s. */
int orbx_popc_peak(int device, double *gpopc_per_s);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */
