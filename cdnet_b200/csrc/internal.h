// internal.h -- launchers shared between the translation units of libcdnet_b200.
#pragma once
#include "common.cuh"

namespace cdnet {

// ddm.cu
int ddm_codes_launch(const uint8_t* cls_maps, uint16_t* codes, uint32_t* flags, int B, int T, int H, int W,
                     int n_classes, cudaStream_t st, int row_lo = 0, int row_hi = -1);

// ccl.cu ------------------------------------------------------------------------------------------
// labels of the connected components of `mask` (non-zero = foreground), raster-first ids 1..n.
// L, idmap: int32 [B,H,W] scratch; rowcnt: int32 [B,H] scratch.  conn 4 or 8.
int ccl_label_launch(const uint8_t* mask, int32_t* labels, int32_t* n_out, int32_t* L, int32_t* idmap,
                     int32_t* rowcnt, int B, int H, int W, int conn, cudaStream_t st);
// skimage.measure.label of a multi-valued uint8 image: 8-connected components of equal non-zero value
int ccl_label_values_launch(const uint8_t* ids, int32_t* labels, int32_t* n_out, int32_t* L, int32_t* idmap,
                            int32_t* rowcnt, int B, int H, int W, cudaStream_t st);
// forest of the 4-connected components of `mask` only (L[p] = root = first raster pixel; background L[p] = p)
int ccl_forest_launch(const uint8_t* mask, int32_t* L, int B, int H, int W, int conn, cudaStream_t st);
// state[p] in {0 bg, 1 fg, 2 filled hole} from a 0/1 mask (scipy binary_fill_holes); leaves in L the
// flattened forest of the equal-value 4-connected components of `mask`.  touch: int32 [B,H,W] scratch.
int fill_holes_state_launch(const uint8_t* mask, uint8_t* state, int32_t* L, int32_t* touch, int B, int H, int W,
                            cudaStream_t st);
// the fused chain fill holes -> remove small (4-conn) -> 8-conn label -> int32 labels (test_dam.py:546-561)
size_t fill_remove_label_workspace(int B, int H, int W);
int fill_remove_label_launch(const uint8_t* inside, int32_t* labels, uint8_t* pred2_out, int B, int H, int W,
                             int min_area, void* ws, size_t ws_bytes, cudaStream_t st);
// per-value pixel counts -> zero labels with count < min_size (integer remove_small_objects)
int remove_small_labels_launch(int32_t* labels, int32_t* counts, int B, int H, int W, int min_size, cudaStream_t st);

// exclusive scan of rowcnt [B,H] over the rows of each tile, in place; n_out[b] = total (may be null)
int scan_rows_launch(int32_t* rowcnt, int32_t* n_out, int B, int H, cudaStream_t st);

// rle.cu: the run-based form of fill holes -> remove small -> 8-connected labels -> dilation by disk(radius <= 2)
size_t rle_tail_workspace(int B, int H, int W);
bool rle_tail_supported(int radius);
int rle_tail_launch(const uint8_t* inside, void* out, int out_elem_bytes, int B, int H, int W, int min_area, int radius,
                    void* ws, size_t ws_bytes, cudaStream_t st);
// plain 4-connected labelling with raster-first ids through the run-based kernels (workspace: rle_tail_workspace)
int rle_label4_launch(const uint8_t* mask, int32_t* labels, int B, int H, int W, void* ws, size_t ws_bytes, cudaStream_t st);
// label4(binary_erosion(binary_fill_holes(marker0))) (postproc_other.py:42-44) in the bit domain; tiles up to 1024 columns
bool rle_markers_supported(int W);
int rle_markers_launch(const uint8_t* marker0, int32_t* labels, int B, int H, int W, void* ws, size_t ws_bytes, cudaStream_t st);

// morph.cu ----------------------------------------------------------------------------------------
int label_dilate_launch(const int32_t* labels, void* out, int out_elem_bytes, int B, int H, int W, int radius,
                        cudaStream_t st);

// edt.cu / watershed.cu
size_t ws_process_workspace(int B, int H, int W);
// marker_rowmax (may be null): int32 [B, H], largest marker id per row before small markers are dropped; with it, rows
// [own_lo, own_hi) are the caller's own rows of an extended tile (CDNET_S_SHARD_OVERFLOW, watershed.cu)
int ws_process_launch(const uint8_t* pred01, int32_t* labels, int32_t* status, int B, int H, int W, int min_size,
                      int ws_flag, void* ws, size_t ws_bytes, cudaStream_t st, int32_t* marker_rowmax = nullptr,
                      int own_lo = 0, int own_hi = 0);
// rowflag: int32 [B * H] scratch (rows with pixels left for the far pass)
int edt_launch(const uint8_t* mask, int32_t* d2, int32_t* g2, int32_t* rowflag, int B, int H, int W, cudaStream_t st);
// squared distance reported where a tile has no background pixel at all (edt.cu)
constexpr int kEdtInf = 1 << 30;

}  // namespace cdnet
