/* superslam_b200 — C ABI of the B200-native SuperPoint + LightGlue stereo front-end.
 *
 * This is the drop-in boundary behind SuperSLAM's inference interfaces
 * (reference include/InferenceInterfaces.h:27-59).  Every entry point takes plain pointers and sizes,
 * returns an int status (0 = ok) and never throws; ssb_last_error() describes the last failure on the
 * calling thread.  The C++ adapter include/superslam_b200_adapter.hpp wraps these calls into
 * superslam::IFeatureExtractor / superslam::IFeatureMatcher; INTEGRATION.md shows the wiring.
 *
 * Conventions shared with the reference:
 *   - images: 8-bit gray (channels = 1) or BGR (channels = 3), row-major, `row_stride` bytes per row
 *   - keypoints: pixel coordinates (x, y) float32, x = w * (W / score_W)   (src/SuperPoint.cc:708-716)
 *   - descriptors: fp16 [count, 256] row-major in DEVICE memory, L2-normalised rows
 *     (include/DescriptorPool.h:13-20, src/DescriptorGather.cu:14-56)
 *   - matches0[i] = index into set 1 or -1, mscores0[i] float32         (src/LightGlue.cc:326-363)
 */
#ifndef SUPERSLAM_B200_H_
#define SUPERSLAM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_OK 0
#define SSB_ERR_INVALID 1
#define SSB_ERR_CUDA 2
#define SSB_ERR_IO 3
#define SSB_ERR_EXHAUSTED 4 /* descriptor pool has no free slot (src/SuperPoint.cc:724-727) */
#define SSB_ERR_NODEVICE 5

typedef struct ssb_superpoint ssb_superpoint;
typedef struct ssb_lightglue ssb_lightglue;
typedef struct ssb_frontend ssb_frontend;
typedef struct ssb_eigenplaces ssb_eigenplaces;

const char* ssb_last_error(void);
int ssb_version(void);
/* Number of sm_100 devices visible to the process, or a negative status. */
int ssb_device_count(void);

/* ---- SuperPoint: replaces class SuperPoint (include/SuperPoint.h:37-52) ------------------------ */

/* SuperPoint(engine_file, max_keypoints, keypoint_threshold, remove_borders) + initialize()
 * (include/SuperPoint.h:39-44, src/SuperPoint.cc:41-63).  `weights_path` is an SSBW archive made by
 * tools/convert_superpoint_weights.py.  `num_slots` <= 0 selects the reference's 8 descriptor slots
 * (include/SuperPoint.h:86-87). */
int ssb_sp_create(const char* weights_path, int max_keypoints, double keypoint_threshold,
                  int remove_borders, int num_slots, int device_id, ssb_superpoint** out);
void ssb_sp_destroy(ssb_superpoint* sp);

/* IFeatureExtractor::extract (batch = 1, src/SuperPoint.cc:894-899) and ::extract_stereo (batch = 2,
 * src/SuperPoint.cc:901-907, one batched {2,1,H,W} pass); larger batches are a throughput extension.
 * For image i: xy[i] receives count[i] (x, y) pairs, score[i] the responses (arrays sized
 * max_keypoints), desc_dev[i] the device pointer of the fp16 [count, 256] rows, slot[i] the pool slot
 * that owns them (-1 when the pool is exhausted: SSB_ERR_EXHAUSTED is returned and that image has no
 * descriptors, like the reference's empty handle).  A slot starts with one reference. */
int ssb_sp_extract(ssb_superpoint* sp, const uint8_t* const* images, int batch, int height, int width,
                   int row_stride, int channels, float* const* xy, float* const* score, int* count,
                   void** desc_dev, int* slot);

/* DeviceDescriptors::slot_ref (include/DescriptorPool.h:62-76): copies share the slot, the last
 * release returns it to the LIFO free list. */
int ssb_sp_slot_retain(ssb_superpoint* sp, int slot);
int ssb_sp_slot_release(ssb_superpoint* sp, int slot);
int ssb_sp_slots_in_use(ssb_superpoint* sp); /* DescriptorPool::in_use() */
int ssb_sp_max_keypoints(ssb_superpoint* sp);
/* Test hook: copy an intermediate device buffer ("conv1a" ... "scores", "grid") to the host. */
int ssb_sp_debug_read(ssb_superpoint* sp, const char* what, void* dst, size_t bytes);

/* ---- LightGlue: replaces class LightGlue (include/LightGlue.h:33-57) --------------------------- */

/* LightGlue(engine_file, image_width, image_height) + initialize() (include/LightGlue.h:36).
 * `weights_path`: SSBW archive with cvg/LightGlue state-dict names (tools/convert_lightglue_weights.py).
 * `max_keypoints` sizes the workspace (the reference engine profile tops out at 1024). */
int ssb_lg_create(const char* weights_path, int image_width, int image_height, int max_keypoints,
                  int device_id, ssb_lightglue** out);
/* LightGlue(shared_engine, w, h) (include/LightGlue.h:39-44, src/SuperSLAM.cc:129-133): shares the
 * immutable weights, owns its stream and workspace; use one context per thread. */
int ssb_lg_clone_context(ssb_lightglue* src, int image_width, int image_height, ssb_lightglue** out);
void ssb_lg_destroy(ssb_lightglue* lg);

/* IFeatureMatcher::match(kp0, DeviceDescriptors, kp1, DeviceDescriptors) (src/LightGlue.cc:377-457).
 * xy*: host pixel keypoints [n, 2]; desc*_dev: device fp16 [n, 256].  matches0 / mscores0: host
 * arrays of n0 entries.  n0 == 0 or n1 == 0 yields no matches and SSB_OK. */
int ssb_lg_match_device(ssb_lightglue* lg, const float* xy0, int n0, const void* desc0_dev,
                        const float* xy1, int n1, const void* desc1_dev, int32_t* matches0,
                        float* mscores0);
/* IFeatureMatcher::match(kp0, cv::Mat, kp1, cv::Mat) (src/LightGlue.cc:285-324): host fp32 [n, 256]. */
int ssb_lg_match_host(ssb_lightglue* lg, const float* xy0, int n0, const float* desc0_f32,
                      const float* xy1, int n1, const float* desc1_f32, int32_t* matches0,
                      float* mscores0);
/* IFeatureMatcher::descriptors_to_host (src/LightGlue.cc:460-475): blocking D2H, fp16 -> fp32. */
int ssb_desc_to_host_f32(int device_id, const void* desc_dev_f16, int count, int dim, float* out);
int ssb_lg_debug_read(ssb_lightglue* lg, const char* what, void* dst, size_t bytes);

/* ---- Frame-pair front end (throughput path; not in the reference API) --------------------------
 * SuperPoint x2 + LightGlue + the StereoFrontEnd disparity / row filter (src/StereoFrontEnd.cc:35-47)
 * for `pairs` stereo pairs per call, chained on one stream with device-resident keypoint counts: one
 * host synchronisation per call.  Image 2p is the left, 2p+1 the right image of pair p. */
int ssb_fe_create(const char* sp_weights, const char* lg_weights, int max_keypoints,
                  double keypoint_threshold, int remove_borders, int lg_image_width, int lg_image_height,
                  float min_disparity, int max_pairs, int device_id, ssb_frontend** out);
void ssb_fe_destroy(ssb_frontend* fe);
/* Host images in (gray u8, 2*pairs pointers), host results out.  Any output pointer may be NULL.
 *   count      [2*pairs]            keypoints per image
 *   xy         [2*pairs][K][2]      score [2*pairs][K]
 *   matches0   [pairs][K]           mscores0 [pairs][K]
 *   stereo_ur  [pairs][K]           right-image u of the accepted match, NaN otherwise
 *   has_depth  [pairs][K]           1 where the match passed the disparity and row checks */
int ssb_fe_process(ssb_frontend* fe, const uint8_t* const* images, int pairs, int height, int width,
                   int row_stride, int* count, float* xy, float* score, int32_t* matches0,
                   float* mscores0, float* stereo_ur, uint8_t* has_depth);
/* Same pipeline on images already resident on the device ([2*pairs][H][W] u8 contiguous); enqueue
 * only, results stay on the device until ssb_fe_fetch.  Used for the HBM-resident throughput number. */
int ssb_fe_enqueue_device(ssb_frontend* fe, const uint8_t* images_dev, int pairs, int height, int width);
int ssb_fe_fetch(ssb_frontend* fe, int pairs, int* count, float* xy, float* score, int32_t* matches0,
                 float* mscores0, float* stereo_ur, uint8_t* has_depth);
/* Streaming form of ssb_fe_process for a sequence of steps: submit enqueues the upload (on a copy stream,
 * into one of two device image buffers), the pipeline and the read-back of one step and returns at once;
 * collect blocks until the OLDEST submitted step has finished and copies its results out (same arrays as
 * ssb_fe_process; `pairs` receives that step's pair count).  At most two steps may be in flight, so the
 * upload of step i+1 overlaps the kernels of step i.  Images must stay valid until their step is collected. */
int ssb_fe_submit(ssb_frontend* fe, const uint8_t* const* images, int pairs, int height, int width,
                  int row_stride);
int ssb_fe_collect(ssb_frontend* fe, int* pairs, int* count, float* xy, float* score, int32_t* matches0,
                   float* mscores0, float* stereo_ur, uint8_t* has_depth);
int ssb_fe_sync(ssb_frontend* fe);
/* Extract-only mode (BASELINE config C1, tests/test_superpoint_only.cc of the reference: SuperPoint alone): the
 * 2*pairs images of a call are independent mono frames; count / xy / score are delivered, the match and stereo
 * outputs are undefined.  Switch only while no streamed step is in flight. */
int ssb_fe_set_extract_only(ssb_frontend* fe, int on);

/* ---- Tracking chain: the SECOND LightGlue call the live pipeline makes per frame -------------------
 * VoEstimator::track matches the last keyframe's left features against the current left features
 * (src/VoEstimator.cc:240-246: matcher_->match(last_keyframe_.keypoints_left, last_keyframe_.descriptors_left,
 * frame.keypoints_left, frame.descriptors_left)) and later makes the frame the new keyframe (:327).  With tracking
 * enabled every pair slot p of a call is a stream that retains its keyframe (keypoints, fp16 descriptor rows, depth
 * flags) on the device, and each call runs SuperPoint x2 + LightGlue (stereo) + post-filter + LightGlue (keyframe <->
 * left) + the depth test of :250-256 in the same captured graph: no descriptor ever visits the host between the two
 * matches.  A stream without a keyframe (before its first promotion) yields no tracking matches, like the reference's
 * first frame (:206-236). */
int ssb_fe_enable_tracking(ssb_frontend* fe, int enable);
int ssb_fe_reset_tracking(ssb_frontend* fe); /* forget every keyframe */
/* "last_keyframe_ = frame" for the streams with promote[p] != 0 (NULL: all `pairs` streams): the left image of pair
 * p of the step delivered last becomes stream p's keyframe, device to device.  The decision is the host's
 * (should_insert_keyframe, src/VoEstimator.cc:291-297); call it after collecting that step, before the next one. */
int ssb_fe_promote_keyframes(ssb_frontend* fe, const uint8_t* promote, int pairs);
/* Tracking outputs of the step delivered last by ssb_fe_process / ssb_fe_fetch / ssb_fe_collect (no further device
 * work): keyframe_count [pairs] = features of the keyframe (0: none yet); track_matches0 [pairs][K]: index into the
 * current LEFT keypoints per keyframe feature (queryIdx = keyframe, trainIdx = frame) or -1; track_mscores0;
 * track_usable [pairs][K] = 1 where the match exists and both ends have stereo depth (:253-256).  Any may be NULL. */
int ssb_fe_tracking_results(ssb_frontend* fe, int pairs, int* keyframe_count, int32_t* track_matches0,
                            float* track_mscores0, uint8_t* track_usable);

/* ---- Multi-device driver: the front end on every GPU of a box, one process -------------------------
 * The object graph of src/SuperSLAM.cc:82-86,107 once per device: one host thread + one ssb_frontend per entry of
 * `device_ids` (NULL: devices 0 .. n_devices-1), weights replicated.  Pair p of a call runs on device p mod n
 * (ssb_mg_device_of_pair) in steps of at most `max_pairs_per_device` pairs through ssb_fe_submit / ssb_fe_collect;
 * there is no data-path collective - every worker writes its pairs' rows into the caller's arrays, which are laid out
 * and indexed exactly like ssb_fe_process's (any may be NULL).  `pairs` is unbounded.  Blocking; not re-entrant. */
typedef struct ssb_multigpu ssb_multigpu;
int ssb_mg_create(const char* sp_weights, const char* lg_weights, int max_keypoints, double keypoint_threshold,
                  int remove_borders, int lg_image_width, int lg_image_height, float min_disparity,
                  int max_pairs_per_device, const int* device_ids, int n_devices, ssb_multigpu** out);
void ssb_mg_destroy(ssb_multigpu* mg);
int ssb_mg_device_count(ssb_multigpu* mg);
int ssb_mg_device_of_pair(ssb_multigpu* mg, int pair);
const char* ssb_mg_last_error(ssb_multigpu* mg); /* of the last failed ssb_mg_process */
int ssb_mg_process(ssb_multigpu* mg, const uint8_t* const* images, int pairs, int height, int width, int row_stride,
                   int* count, float* xy, float* score, int32_t* matches0, float* mscores0, float* stereo_ur,
                   uint8_t* has_depth);
/* CUDA-event timing on the front end's own stream (torch.cuda.Event only sees torch's stream). */
int ssb_fe_event_record(ssb_frontend* fe, int index /* 0..15 */);
int ssb_fe_event_elapsed_ms(ssb_frontend* fe, int start_index, int stop_index, float* ms);
/* Device buffer helpers for callers without a CUDA runtime binding of their own. */
int ssb_fe_upload_images(ssb_frontend* fe, const uint8_t* const* images, int count, int height, int width,
                         int row_stride, uint8_t** images_dev_out);
int ssb_fe_kernel_launches_per_call(ssb_frontend* fe, int pairs);
/* Process-wide instrumentation used by bench.py: kernels launched so far, and optional CUDA-event
 * timing of every kernel on its launching stream (enable, run steps, collect after each sync, report as
 * "label count total_ms" lines). */
long long ssb_kernel_launch_count(void);
void ssb_profile_enable(int on);
void ssb_profile_collect(void);
int ssb_profile_report(char* buf, size_t bytes);
ssb_superpoint* ssb_fe_superpoint(ssb_frontend* fe);
ssb_lightglue* ssb_fe_lightglue(ssb_frontend* fe);

/* ---- Image front door: the data formats either side of the networks ---------------- */
typedef struct ssb_rectifier ssb_rectifier;
typedef struct ssb_rgbd ssb_rgbd;

/* cv::remap(img, out, M1, M2, cv::INTER_LINEAR) with the CV_32FC1 maps of cv::initUndistortRectifyMap and the
 * default constant-0 border: the EuRoC rectification the reference does on the host before track_stereo
 * (examples/stereo/euroc.cc:118-133,176-177).  map_x / map_y: float32 [dst_height][dst_width]; results are
 * bit-identical to OpenCV's fixed-point bilinear path.  dst_height * dst_width must be a multiple of 4. */
int ssb_rect_create(const float* map_x, const float* map_y, int dst_height, int dst_width, int src_height,
                    int src_width, int max_images, int device_id, ssb_rectifier** out);
void ssb_rect_destroy(ssb_rectifier* r);
/* The host half of ssb_rect_create on its own (no device needed): OpenCV's conversion of `count` CV_32F map entries
 * to fixed point - xy[i] = (uint16)ix | (uint16)iy << 16 (int16 each, saturated), frac[i] = fy * 32 + fx. */
int ssb_rect_convert_maps(const float* map_x, const float* map_y, size_t count, uint32_t* xy, uint16_t* frac);
/* host gray u8 images in (`row_stride` bytes per row), host images out (dst_width bytes per row) */
int ssb_rect_remap(ssb_rectifier* r, const uint8_t* const* images, int count, int row_stride,
                   uint8_t* const* out);
/* device [count][src_h][src_w] -> device [count][dst_h][dst_w]; enqueued on the rectifier's stream, then
 * synchronised (chain it in front of ssb_fe_enqueue_device to keep rectified images off the host) */
int ssb_rect_remap_device(ssb_rectifier* r, const uint8_t* src_dev, int count, uint8_t* dst_dev);
/* Put a pair of rectifiers in front of the frame-pair pipeline: every ssb_fe_process / submit / enqueue_device
 * call then takes RAW images (the rectifiers' source size), remaps image 2p with `left` and 2p+1 with `right` on
 * the device and runs SuperPoint x2 + LightGlue on the rectified pair - the EuRoC loop of
 * examples/stereo/euroc.cc:176-181 without the host remap.  The rectifiers are borrowed (they must outlive the
 * front end or be unset with NULL, NULL) and must agree on sizes. */
int ssb_fe_set_rectifiers(ssb_frontend* fe, ssb_rectifier* left, ssb_rectifier* right);

/* RgbdFrontEnd::process after the extraction (include/RgbdFrontEnd.h:15-37, src/RgbdFrontEnd.cc:24-58):
 * cv::undistortPoints(raw, undist, K, D, noArray(), K) when dist has a non-zero entry, depth sampled at
 * lround(raw), stereo = (uL, uL - bf / Z, v) and has_depth = 1 iff 0 < Z < max_depth, else (uL, NaN, v). */
int ssb_rgbd_create(int max_keypoints, int max_height, int max_width, int device_id, ssb_rgbd** out);
void ssb_rgbd_destroy(ssb_rgbd* r);
/* xy: host float [n][2] raw keypoints (ssb_sp_extract output).  depth: host image, depth_type 0 = CV_16U,
 * 1 = CV_32F, `row_stride` bytes per row.  camera = {fx, fy, cx, cy}; dist: n_dist (<= 14) coefficients in
 * OpenCV order or NULL.  out_xy float [n][2], out_stereo double [n][3], out_has_depth uint8 [n]. */
int ssb_rgbd_process(ssb_rgbd* r, const float* xy, int n, const void* depth, int depth_type, int height,
                     int width, int row_stride, const double* camera, const double* dist, int n_dist,
                     double bf, double depth_factor, double max_depth, float* out_xy, double* out_stereo,
                     uint8_t* out_has_depth);

/* ---- EigenPlaces: replaces class EigenPlaces (include/EigenPlaces.h:21-66) ----------------------
 * superslam::IPlaceRecognizer (include/PlaceRecognizer.h:20-36) over a B200-native ResNet18 + GeM + FC
 * global-descriptor network and a device-resident CosineDescriptorIndex (src/PlaceRecognizer.cc:21-52). */

/* EigenPlaces(engine_file, input_width, input_height) + initialize() (include/EigenPlaces.h:26-30).
 * `weights_path`: SSBW archive with the state-dict names of the torch.hub model
 * (utils/convert_eigenplaces_to_onnx.py:54-60; superslam_b200/eigenplaces_weights.py).  Both sides of the
 * network input must be multiples of 32.  `max_batch` images are processed per pass (the reference: 1). */
int ssb_ep_create(const char* weights_path, int input_width, int input_height, int max_batch, int device_id,
                  ssb_eigenplaces** out);
void ssb_ep_destroy(ssb_eigenplaces* ep);
int ssb_ep_descriptor_dim(ssb_eigenplaces* ep); /* 512 */
/* IPlaceRecognizer::compute_global_descriptor (src/EigenPlaces.cc:145-174) for `count` same-size images:
 * gray / BGR u8 -> RGB, cv::resize to the network input, /255, ImageNet normalisation (:123-143), network,
 * final L2 normalisation.  descriptors: host [count][512] float32. */
int ssb_ep_compute(ssb_eigenplaces* ep, const uint8_t* const* images, int count, int height, int width,
                   int row_stride, int channels, float* descriptors);
/* IPlaceRecognizer::add / ::query (include/EigenPlaces.h:36-42, src/PlaceRecognizer.cc:21-52): rows are
 * re-normalised on insertion, insertion order is recency; a query returns at most `top_k` (and `capacity`)
 * entries with score >= min_score among all but the last `exclude_recent` insertions, best first. */
int ssb_ep_add(ssb_eigenplaces* ep, uint64_t keyframe_id, const float* descriptor, int dim);
int ssb_ep_query(ssb_eigenplaces* ep, const float* descriptor, int dim, uint64_t exclude_recent, int top_k,
                 float min_score, uint64_t* keyframe_ids, float* scores, int capacity, int* n_out);
int ssb_ep_index_size(ssb_eigenplaces* ep); /* CosineDescriptorIndex::size() */
int ssb_ep_debug_read(ssb_eigenplaces* ep, const char* what, void* dst, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* SUPERSLAM_B200_H_ */
