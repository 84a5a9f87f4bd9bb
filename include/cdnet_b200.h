/*
 * cdnet_b200.h -- C ABI of libcdnet_b200.so: CDNet's geometry hot path on B200 (sm_100a).
 *
 * The reference (honglianghe/CDNet) has no FFI: its boundary for this path is a set of plain
 * Python callables on numpy arrays.  Each entry point below names the
 * reference callable / line range it replaces; the Python mirror with the reference's own
 * signatures lives in cdnet_b200/api.py and binds these symbols with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (dense, row-major, batch first);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and performs
 *     no host synchronisation and no allocation: scratch comes from the caller's workspace `ws`
 *     whose size is returned by the matching *_workspace_bytes() (0 on invalid arguments);
 *   - return value: 0 = ok, <0 = -(cudaError_t), >0 = CDNET_E_* domain error; nothing throws;
 *   - B tiles of H x W pixels; H*W < 2^31; pixel p = y*W + x; tiles are independent.
 */
#ifndef CDNET_B200_H_
#define CDNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDNET_E_BADARG 1    /* null pointer, unsupported class count / connectivity / radius ... */
#define CDNET_E_WORKSPACE 2 /* workspace smaller than *_workspace_bytes() */

/* per-tile status bits written by the kernels into `status[b]` (int32, device) */
#define CDNET_S_DDM_CONSTANT 1 /* a direction map was constant: generate_dd_map's 0/0 -> NaN
                                  (getDirectionDiffMap.py:104-106); the reference's
                                  `assert(np.min(enhanced_boundary) >= 0)` (test_dam.py:535) fails */
#define CDNET_S_WS_OVERFLOW 2  /* watershed queue overflow (cannot happen with the sizes from
                                  *_workspace_bytes; kept as a guard) */

#define CDNET_S_NO_BACKGROUND 16 /* watershed path: the mask has no background pixel; the reference's
                                  * gen_inst_dst_map raises ValueError there (`nuc_list.remove(0)`,
                                  * postproc_other.py:18-19) */
#define CDNET_S_WS_CONTESTED_SHIFT 8 /* watershed path: status >> 8 = number of mask pixels that two age-0 markers of
                                      * EQUAL priority and different labels compete for (their pop order is where
                                      * scikit-image's heap and this build's raster order may differ; 0 = the result
                                      * does not depend on that order to first order).  Flags live in bits 0..7. */
#define CDNET_S_SHARD_OVERFLOW 64 /* cdnet_shard_ws_process: a nucleus reaches beyond the overlap rows of the tile */
#define CDNET_S_CLASS_RANGE 32   /* cdnet_direction_one_hot: a class id outside [0, C); the reference's
                                  * `target_direction_temp[j, k]` raises IndexError (train_util_dam.py:138) */
#define CDNET_S_PAIR_OVERFLOW 4 /* cdnet_label_pairs: more distinct (true, pred) pairs than `cap` */
#define CDNET_S_PAIR_RANGE 8    /* cdnet_label_pairs: a label id is negative or above INT32_MAX */

const char* cdnet_version(void);
/* 1 if the library was built for the device `device` can run (compute capability 10.x) */
int cdnet_device_ok(int device);

/* ---- generate_dd_map(label_direction, direction_classes) ------------------------------------
 * data_prepare/getDirectionDiffMap.py:44-108 (with circshift :14-42 and
 * DTOffsetHelper.label_to_vector, data_prepare/SegFix_offset_helper.py:246-261).
 * cls: uint8 [B,H,W] class maps; out: float32 [B,H,W] in {0, .5, 1} (NaN for a constant map,
 * status bit CDNET_S_DDM_CONSTANT).  n_classes in {5, 9, 17}.  status may be NULL. */
size_t cdnet_ddm_workspace_bytes(int B, int H, int W);
int cdnet_ddm(const uint8_t* cls, float* out, int32_t* status, int B, int H, int W, int n_classes,
              void* ws, size_t ws_bytes, void* stream);

/* ---- circshift(matrix_ori, direction, shiftnum1, shiftnum2) ---------------------------------
 * data_prepare/getDirectionDiffMap.py:14-42: zero-filled shift of C planes of `elem_bytes`
 * (1, 2, 4 or 8) byte elements.  direction 1..4. */
int cdnet_circshift(const void* in, void* out, int C, int H, int W, int elem_bytes, int direction,
                    int shift1, int shift2, void* stream);

/* ---- connected-component labelling -----------------------------------------------------------
 * scipy.ndimage.label (4-connected, postproc_other.py:37,44) / skimage.measure.label
 * (8-connected, test_dam.py:561, test.py:292, my_transforms_direction.py:755,773): non-zero =
 * foreground, ids 1..n in raster order of each component's first pixel.
 * mask: uint8 [B,H,W]; labels: int32 [B,H,W]; n_out: int32 [B] component counts (may be NULL). */
size_t cdnet_ccl_workspace_bytes(int B, int H, int W);
int cdnet_ccl(const uint8_t* mask, int32_t* labels, int32_t* n_out, int B, int H, int W,
              int connectivity, void* ws, size_t ws_bytes, void* stream);

/* ---- skimage.measure.label of a multi-valued image (my_transforms_direction.py:723-725) ----------
 * ids: uint8 [B,H,W]; labels: int32 [B,H,W], 8-connected components of EQUAL non-zero value, numbered 1..n in
 * raster order of their first pixel; n_out (may be NULL): int32 [B].  Workspace: cdnet_ccl_workspace_bytes. */
int cdnet_label_values(const uint8_t* ids, int32_t* labels, int32_t* n_out, int B, int H, int W, void* ws,
                       size_t ws_bytes, void* stream);

/* ---- scipy.ndimage.binary_fill_holes (test_dam.py:546, test.py:277, postproc_other.py:42,51) --
 * out may alias mask. */
size_t cdnet_fill_holes_workspace_bytes(int B, int H, int W);
int cdnet_fill_holes(const uint8_t* mask, uint8_t* out, int B, int H, int W, void* ws,
                     size_t ws_bytes, void* stream);

/* ---- skimage.morphology.remove_small_objects --------------------------------------------------
 * bool input (test_dam.py:548, test.py:279): 4-connected components smaller than min_size are
 * cleared.  out may alias mask. */
size_t cdnet_remove_small_mask_workspace_bytes(int B, int H, int W);
int cdnet_remove_small_mask(const uint8_t* mask, uint8_t* out, int B, int H, int W, int min_size,
                            void* ws, size_t ws_bytes, void* stream);
/* integer input (postproc_other.py:46,48,53): labels whose pixel count is < min_size are zeroed,
 * no renumbering.  Label values must lie in [0, H*W].  In place. */
size_t cdnet_remove_small_labels_workspace_bytes(int B, int H, int W);
int cdnet_remove_small_labels(int32_t* labels, int B, int H, int W, int min_size, void* ws,
                              size_t ws_bytes, void* stream);

/* ---- skimage.morphology.dilation(labels, disk(radius)) ----------------------------------------
 * test_dam.py:563, test.py:295, my_transforms_direction.py:760,774,819: max over the disk
 * footprint (radius 1 = 5-px cross, 2 = 13 px), out-of-image taps ignored.
 * out_elem_bytes 4 (int32) or 8 (int64, what measure.label hands on). */
int cdnet_label_dilate(const int32_t* labels, void* out, int out_elem_bytes, int B, int H, int W,
                       int radius, void* stream);

/* ---- scipy.ndimage.distance_transform_edt ------------------------------------------------------
 * postproc_other.py:24, my_transforms_direction.py:802,822.  d2: exact squared distance (int32)
 * to the nearest zero pixel of mask; dist (may be NULL): float64 sqrt of it. */
size_t cdnet_edt_workspace_bytes(int B, int H, int W);
int cdnet_edt(const uint8_t* mask, int32_t* d2, double* dist, int B, int H, int W, void* ws,
              size_t ws_bytes, void* stream);

/* ---- postproc_other.process(pred, model_mode, min_size, ws) -----------------------------------
 * postproc_other.py:15-54.  pred01: uint8 [B,H,W], already binarised (pred > 0.5, :31-32).
 * ws != 0: label -> per-instance EDT scaled to uint8 -> markers (>125, fill holes, cross
 * erosion, label, remove small) -> marker-controlled watershed on the uint8-negated distance,
 * (value, age, raster index) order -> remove small.  ws == 0 (model_mode unet/micronet, :35,
 * :50-54): fill holes -> label -> remove small.  labels: int32 [B,H,W], ids keep gaps. */
size_t cdnet_ws_postproc_workspace_bytes(int B, int H, int W);
int cdnet_ws_postproc(const uint8_t* pred01, int32_t* labels, int32_t* status, int B, int H, int W,
                      int min_size, int ws_flag, void* ws, size_t ws_bytes, void* stream);
/* process() (ws = 1) on the extended tile [H, W] of one rank of a row-sharded slide (cdnet_b200/sharded.py,
 * postproc = 1): rows [own_lo, own_hi) are the rank's own, the rows around them overlap the neighbours' so that every
 * nucleus touching the own rows is seen whole.  marker_rowmax int32 [H] receives the largest marker id per row as
 * postproc_other.py:44 numbers them -- before remove_small_objects (:46), whose gaps stay -- so that
 * max(marker_rowmax[0..y]) = markers that start on rows 0..y; status (required) also gets CDNET_S_SHARD_OVERFLOW when
 * a 4-connected component of the mask touches both the own rows and an outer overlap row (overlap too small). */
int cdnet_shard_ws_process(const uint8_t* pred01, int32_t* labels, int32_t* marker_rowmax, int32_t* status, int H, int W,
                           int own_lo, int own_hi, int min_size, void* ws, size_t ws_bytes, void* stream);
/* Own rows of that tile: tile-local marker ids -> slide-global ids.  scalars int32 [3] on the device = {markers that
 * start above the own rows, markers that start on them, markers owned by lower ranks}; lut int32 [> largest local id]
 * holds the ids adopted from the row neighbours (0 = none); err int32 [1] is set to 1 (never cleared) when a label has
 * neither.  labels / out: int32 [rows, W]. */
int cdnet_shard_ws_relabel(const int32_t* labels, const int32_t* scalars, const int32_t* lut, int32_t* out, int32_t* err,
                           int rows, int W, void* stream);

/* ---- direction-aware inference post-processing, test_dam.py:455-563 ----------------------------
 * dcm:   uint8   [B,n_maps,H,W]  n_maps = 8: the 8 TTA direction-argmax maps (prob_dcm ...
 *                           prob_dcm_r90_hvf, dcm_combined = 1, :455-498); n_maps = 1: the single-map
 *                           variant (:499-502)
 * prob:  float32 [B,3,H,W]  class probabilities; if write_prob != 0 channel 2 is overwritten
 *                           with the boosted boundary probability like the reference (:536)
 * point: float32 [B,1,H,W]  point map
 * out:   labels [B,H,W], int32 or int64 (out_elem_bytes 4 / 8; the reference returns int64
 *        from measure.label when postproc == 0 and int32 from process() when postproc == 1)
 * postproc: 0 = fill holes / remove small / measure.label (:546-561), 1 = postproc_other.process with the
 *        watershed (:559), 2 = process as it runs for model_mode 'unet' (no watershed, postproc_other.py:35,50-54)
 * status: int32 [B] (CDNET_S_*), may be NULL. */
size_t cdnet_dam_postproc_workspace_bytes(int B, int H, int W);
int cdnet_dam_postproc(const uint8_t* dcm, int n_maps, float* prob, const float* point, void* out,
                       int out_elem_bytes, int32_t* status, int B, int H, int W,
                       int direction_classes, int min_area, int radius, int postproc,
                       int write_prob, void* ws, size_t ws_bytes, void* stream);

/* ---- DcmVoting2(direct_map), utils.py:1150-1159 (dead by default in the reference: voting_firt = 0,
 * test_dam.py:471; SURVEY.md section 8f "next") -----------------------------------------------
 * dcm: uint8 [B,8,H,W] the 8 TTA direction maps (classes 0..8); out: uint8 [B,H,W] voted class. */
int cdnet_dcm_voting2(const uint8_t* dcm, uint8_t* out, int B, int H, int W, void* stream);

/* ---- plain inference post-processing, test.py:270-295 -------------------------------------------
 * prob: float32 [B,C,H,W]; multi_class != 0: inside = (argmax == 1) over C channels, else
 * inside = prob[:,0] >= 0.5.  Then as above without the direction map; process() gets
 * min_size = min_area (:289-290). */
size_t cdnet_plain_postproc_workspace_bytes(int B, int H, int W);
int cdnet_plain_postproc(const float* prob, int C, void* out, int out_elem_bytes, int32_t* status,
                         int B, int H, int W, int multi_class, int min_area, int radius,
                         int postproc, void* ws, size_t ws_bytes, void* stream);

/* ---- get_centerpoint2(mask, n, m) for every label at once --------------------------------------
 * my_transforms_direction.py:651-685.  labels: int32 [B,H,W] (ids in [0, max_label]);
 * centres: int32 [B, max_label+1, 2] = (row, col) of the first raster-order pixel of maximum
 * centerness of each id, (-1,-1) for absent ids. */
size_t cdnet_center_points_workspace_bytes(int B, int H, int W, int max_label);
int cdnet_center_points(const int32_t* labels, int32_t* centres, int B, int H, int W, int max_label,
                        void* ws, size_t ws_bytes, void* stream);

/* ---- label statistics for LabelEncoding's branch selection -------------------------------------
 * my_transforms_direction.py:714-719 (`len(np.unique(label_inside)) > 2` = instance-level input).
 * presence: int32 [B,256], 1 where the value occurs; fg_count: int32 [B] non-zero pixels. */
int cdnet_label_stats(const uint8_t* ids, int32_t* presence, int32_t* fg_count, int B, int H, int W,
                      void* stream);

/* ---- LabelEncoding.__call__ (out_c = 3, do_direction = 1), my_transforms_direction.py:697-885 --
 * ids:       uint8 [B,H,W]  channel 0 of the label image (data_folder.py:29,37)
 * instance_level: 1 = ids are instance ids (reference: > 2 unique values, :743-760), 0 = a {0,255} three-class
 *            label (:763-774); the out_c != 3 forms of the transform (:721-739, no boundary class, instances
 *            not dilated, ternary in {0,255}): 2 = instance ids, 3 = {0,255} label with ids = max(channel 0,
 *            channel 1).  4 / 5 = my_transforms.LabelEncoding with do_direction = 1 (my_transforms.py:661-836, out_c = 3)
 *            on instance ids / on a {0,255} label: that module's ternary rule, instances = measure.label of the label
 *            (8-connected, equal value, not dilated, no watershed), nucleus centre = first raster maximum of the
 *            nucleus's own exact distance transform (peak_local_max(..., exclude_border=0, num_peaks=1), :775).
 *            Applies to all B tiles of the call
 * ternary:   uint8 [B,H,W]  {0,127,255}
 * point:     float16 bits [B,H,W] Gaussian point map (sigma 2)
 * direction: int64 [B,H,W]  classes 0..num_classes; num_classes in {8, 16} (the reference's env
 *            dt_num_classes, data_prepare/SegFix_offset_helper.py:37-39)
 * inst_out (may be NULL): int32 [B,H,W] the dilated instance map (label_instance)
 * dir_out  (may be NULL): float32 [B,H,W,2] the composited Sobel direction map (dir_map)
 * gauss_w  (HOST pointer, may be NULL): the 17 float64 weights of
 *            scipy.ndimage gaussian_filter(sigma=2); NULL = computed with libm exp(). */
size_t cdnet_encode_targets_workspace_bytes(int B, int H, int W);
int cdnet_encode_targets(const uint8_t* ids, int instance_level, uint8_t* ternary, uint16_t* point,
                         int64_t* direction, int32_t* inst_out, float* dir_out, int32_t* status,
                         int B, int H, int W, int num_classes, const double* gauss_w, void* ws,
                         size_t ws_bytes, void* stream);

/* The same transform on int32 instance ids, i.e. on label images loaded WITHOUT the uint8 truncation of
 * data_folder.py:26,29,37 (`img.astype(np.uint8)` wraps ids at 256, so two touching nuclei whose ids are congruent
 * mod 256 lose the boundary between them).  out_c = 3 forms only: instance_level 1 (instance ids) or 0 ({0,255} label).
 * cdnet_label_stats_i32: n_distinct[b] = min(number of distinct values of tile b, 3) -- the reference only asks
 * `len(np.unique(label_inside)) > 2` (my_transforms_direction.py:714-719) --, fg_count[b] = non-zero pixels;
 * scratch: int32 [B,5]. */
int cdnet_encode_targets_i32(const int32_t* ids, int instance_level, uint8_t* ternary, uint16_t* point,
                             int64_t* direction, int32_t* inst_out, float* dir_out, int32_t* status,
                             int B, int H, int W, int num_classes, const double* gauss_w, void* ws,
                             size_t ws_bytes, void* stream);
int cdnet_label_stats_i32(const int32_t* ids, int32_t* n_distinct, int32_t* fg_count, int32_t* scratch, int B, int H,
                          int W, void* stream);

/* ---- device-resident hand-off from the CNN ---------------------------
 * test_dam.py:299-450 + get_probmaps :983-1013 in one kernel: for each of the 8 test-time-augmentation
 * variants (order: identity, hf, vf, hvf, r90, r90_hf, r90_vf, r90_hvf; the rotated ones live in a [W,H] frame)
 * float32 softmax of the 3-class head, float32 softmax of the direction head, direction[0] *= mask[0],
 * first-maximum argmax; np.flip / np.rot90(k=3) back to the original frame; probabilities and point maps summed
 * in variant order and divided by 8.
 * mask_logits / point / dir_logits: HOST arrays of n_variants DEVICE pointers: float32 [B,3,h_v,w_v],
 * [B,1,h_v,w_v], [B,dir_classes,h_v,w_v]; prob_out float32 [B,3,H,W]; point_out float32 [B,1,H,W]; dcm_out
 * uint8 [B,n_variants,H,W] -- the three inputs of cdnet_dam_postproc.  n_variants 8, or 1 = no augmentation
 * (`tta` off, test_dam.py:313: the identity variant alone, nothing averaged).  dir_classes in {5, 9, 17}. */
int cdnet_tta_merge(const float* const* mask_logits, const float* const* point,
                    const float* const* dir_logits, int n_variants, float* prob_out, float* point_out,
                    uint8_t* dcm_out, int B, int H, int W, int dir_classes, void* stream);

/* ---- the direction quantiser as stand-alone operators --------------------------------------------
 * DTOffsetHelper static methods of data_prepare/SegFix_offset_helper.py; `n` counts elements.
 *
 * cdnet_label_to_vector  (:246-261, table :50-89): labels [N, plane] (uint8 / int32 / int64 by
 *   elem_bytes) -> int64 [N, 2, plane] = (dh, dw); ids the table does not hold give (0, 0).
 *   num_classes in {4, 5, 8, 9, 16, 17, 32} (c4_align_axis unset).
 * cdnet_align_angle      (:311-341; num_classes 4 -> align_angle_c4 :286-309): angle (float32 / float64 by
 *   in_elem_bytes) -> snapped angle (float64 on the reference's numpy path, float32 on its torch path and
 *   for 4 classes; may be NULL) and int64 bin index (may be NULL).  num_classes in {4, 8, 16, 32}.
 * cdnet_angle_to_vector  (:423-450): angle -> snapped angle -> (sin, cos) [n, 2].  `table` is a HOST pointer to
 *   (num_classes + 1) x 2 float64: (sin, cos) of every bin centre as the caller's numpy / torch evaluates
 *   them, then (sin 0, cos 0) for angles no bin matches (NaN).
 * cdnet_vector_to_label  (:486-506): (v0, v1) [n, 2] float32 / float64 -> int64 bin index of
 *   degrees(atan2(v0, v1)). */
int cdnet_label_to_vector(const void* labels, int elem_bytes, int64_t* out, int N, size_t plane,
                          int num_classes, void* stream);
int cdnet_align_angle(const void* angle, int in_elem_bytes, void* snapped, int out_elem_bytes,
                      int64_t* index, size_t n, int num_classes, void* stream);
int cdnet_angle_to_vector(const void* angle, int in_elem_bytes, void* vec, int out_elem_bytes,
                          const double* table, size_t n, int num_classes, void* stream);
int cdnet_vector_to_label(const void* vec, int elem_bytes, int64_t* label, size_t n, int num_classes,
                          void* stream);

/* ---- training-side consumers of the targets ------------------------
 * cdnet_direction_one_hot: train_util_dam.py:123-142.  direction int64 [B, plane] class ids; target0: the
 *   ternary target {0,1,2} of TILE 0 [plane] (uint8 or int64 by target_elem_bytes) -- the reference masks every
 *   tile of the batch with `target[0]` (:139); out float32 [B, C, plane]: channel k = 1 where direction == k on
 *   foreground; a tile with one distinct direction value gets channel 0 = 1 everywhere (:141).
 *   status (may be NULL): CDNET_S_CLASS_RANGE.
 * cdnet_ternary_label: my_transforms.py:661-761, LabelEncoding without direction.  ch0 / ch1 (ch1 only for
 *   mode 3, may be NULL): uint8 [B,H,W] label channels; out uint8 [B,H,W] in {0,127,255}.
 *   mode 0: out_c == 3, instance ids (:713-727); 1: out_c == 3, {0,255} label (:728-742);
 *   mode 2: out_c != 3, instance ids (:690-699); 3: out_c != 3, {0,255} label (:700-708). */
size_t cdnet_direction_one_hot_workspace_bytes(int B);
int cdnet_direction_one_hot(const int64_t* direction, const void* target0, int target_elem_bytes,
                            float* out, int32_t* status, int B, int C, size_t plane, void* ws,
                            size_t ws_bytes, void* stream);
int cdnet_ternary_label(const uint8_t* ch0, const uint8_t* ch1, int mode, uint8_t* out, int B, int H,
                        int W, void* stream);

/* ---- whole-slide row shards ---------------------------------------------
 * One EXTENDED tile per call: the shard's own rows plus one ghost row of the neighbouring shard on
 * each inner side; planes are [He, W].  The host (cdnet_b200/sharded.py) reconciles what straddles
 * the shard seams between the stages: scalar all-reduces of the DDM value flags and the point-map
 * maximum, and a union of the seam components' roots for fill-holes (frame-touch flags),
 * remove-small (areas) and the canonical raster-order numbering.  No reference counterpart: the
 * reference post-processes a slide on one CPU (test_dam.py:455-563). */
int cdnet_shard_ddm_codes(const uint8_t* dcm, uint16_t* codes, uint32_t* flags, int T, int He, int W,
                          int n_classes, int row_lo, int row_hi, void* stream);
int cdnet_shard_point_max(const float* point, uint32_t* pmax, size_t n, void* stream);
int cdnet_shard_boost(const uint16_t* codes, const uint32_t* flags, const float* point,
                      const uint32_t* pmax, float* prob, uint8_t* inside, int32_t* status, int He,
                      int W, int n_maps, int write_prob, void* stream);
/* stage 1: equal-value 4-conn forest (L = root per pixel), frame-touch flags per root */
int cdnet_shard_label_stage1(const uint8_t* inside, int32_t* L, int32_t* touch, int He, int W,
                             int top_is_frame, int bottom_is_frame, void* stream);
/* stage 2: state (0 bg / 1 fg / 2 hole), holes joined, areas over rows [row_lo,row_hi) (area zeroed by caller) */
int cdnet_shard_label_stage2(const uint8_t* inside, int32_t* L, const int32_t* touch, uint8_t* state,
                             int32_t* area, int He, int W, int row_lo, int row_hi, void* stream);
/* stage 3: keep = area >= min_area, diagonal (8-conn) joins, L = root per pixel */
int cdnet_shard_label_stage3(const uint8_t* state, int32_t* L, const int32_t* area, uint8_t* keep,
                             int min_area, int He, int W, void* stream);
/* stage 4: raster-order ids 1..n_owned for kept roots not marked in `excluded` (may be NULL) */
int cdnet_shard_label_stage4(int32_t* L, const uint8_t* keep, const uint8_t* excluded, int32_t* idmap,
                             int32_t* rowcnt, int32_t* n_owned, int He, int W, void* stream);
int cdnet_shard_relabel(const int32_t* L, const uint8_t* keep, const int32_t* idmap, int32_t* labels,
                        int He, int W, void* stream);

/* ---- device-side seam reconciliation for row shards (csrc/seam.cu) ---------------------------------
 * A rank exports the run starts of the rows it shares with its neighbours as a table of int32x4 rows
 * (gid, neighbour gid or -1, attr, attr-once-per-root); row 0 = (count, error bits, 0, 0).  The host
 * all-gathers the tables ([nranks, cap, 4] int32, NCCL) and every rank solves the same union on its
 * device: mode 0 = OR of attr per class -> plane at the local roots, 1 = SUM, 2 = mark the seam roots
 * this rank does not own in `excluded` (then cdnet_seam_ids_export / _apply hand out the class ids).
 * No host synchronisation anywhere.  ws: cdnet_seam_workspace_bytes(nranks, cap), reused across a round. */
size_t cdnet_seam_workspace_bytes(int nranks, int cap);
int cdnet_seam_export(const int32_t* L, const uint8_t* valid, const int32_t* attr, int32_t* emitted,
                      int round_id, int off, int He, int W, int has_top, int has_bottom,
                      const int32_t* nb_gid, int32_t* tbl, int cap, void* stream);
int cdnet_seam_solve(const int32_t* gathered, int nranks, int cap, int my_rank, int mode, int off,
                     int own_lo, int own_hi, int32_t* plane, uint8_t* excluded, void* ws, size_t ws_bytes,
                     void* stream);
int cdnet_seam_ids_export(const int32_t* gathered, int nranks, int cap, int my_rank, int off, int own_lo,
                          int own_hi, const int32_t* idmap, int32_t* emitted, int round_id, int32_t* tbl2,
                          void* ws, size_t ws_bytes, void* stream);
int cdnet_seam_ids_apply(const int32_t* gathered, const int32_t* gathered2, int nranks, int cap,
                         int my_rank, int off, int32_t* idmap, void* ws, size_t ws_bytes, void* stream);

/* ---- instance metrics: pair table of two label images (SURVEY.md section 8f "next", rank 2) --------
 * The one reduction behind stats_utils.py:7-98 get_fast_aji, :101-177 get_fast_aji_plus, :182-275
 * get_fast_pq, :279-318 get_fast_dice_2, :324-334 get_dice_1, :338-357 get_dice_2 and :361-389
 * remap_label (called from test.py:342-345, test_dam.py:616-621): for every tile the distinct pairs
 * (t, q) = (true[p], pred[p]) with their pixel counts (they add up to H * W).  true_lab / pred_lab: int32 or int64
 * [B,H,W] (elem_bytes 4 / 8), ids in [0, INT32_MAX]; keys: uint64 [B,cap] = t << 32 | q; counts: int32
 * [B,cap]; n_out: int32 [B] pairs written (arbitrary order); status: int32 [B], CDNET_S_PAIR_*.
 * Areas are the row / column sums of the table; the float64 epilogue is host code (cdnet_b200/metrics.py). */
size_t cdnet_label_pairs_workspace_bytes(int B, int cap);
int cdnet_label_pairs(const void* true_lab, const void* pred_lab, int elem_bytes, uint64_t* keys,
                      int32_t* counts, int32_t* n_out, int32_t* status, int B, int H, int W, int cap,
                      void* ws, size_t ws_bytes, void* stream);
/* remap_label's gather (stats_utils.py:386-388): out[i] = new_ids[j] where sorted_ids[j] == in[i], else 0.
 * in: int32 / int64 [n]; out: int32 [n]; sorted_ids ascending, new_ids: int32 [n_ids] (device). */
int cdnet_remap_labels(const void* in, int elem_bytes, int32_t* out, const int32_t* sorted_ids,
                       const int32_t* new_ids, int n_ids, size_t n, void* stream);

/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
unsigned long long cdnet_launch_count(void);

/* optional per-launch CUDA-event timing (bench.py's live roofline numbers): enable, run, then
 * cdnet_profile_report() synchronises the device and writes "kernel\tlaunches\ttotal_ms\n" lines
 * into buf; returns the number of distinct kernels and clears the records. */
void cdnet_profile_enable(int on);
int cdnet_profile_report(char* buf, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* CDNET_B200_H_ */
