/*
 * p3d_imageio.h - C ABI of the image output path behind the renderer.
 *
 * The reference's eval sweep ends every view with `I(out['image']).save(fn)` and `I(xyza).save(fn)`
 * (/root/reference/_scripts/eval/generate.py:141-148): a blocking `.cpu()` of fp32 images, torchvision's
 * `to_pil_image` (clamp to [0,1], x255, truncate to uint8: /root/reference/_util/twodee_v1.py:174-185) and PIL's PNG
 * encoder, all on the one Python thread that also drives the GPU.  Here
 *   - quantisation, channel interleave and PNG scan-line filtering (the five PNG filters + libpng's minimum-sum
 *     heuristic) run on the device: a view leaves the GPU as 0.79 MB of filtered bytes instead of 3.1 MB of fp32;
 *   - the copy to pinned host memory is asynchronous on the caller's stream, and a pool of host threads deflates
 *     (zlib) and writes the files, so rank 0's render loop never waits for an encoder.
 * Files are PNGs any decoder reads; the decoded PIXELS equal what the reference's `I(...).save` writes (the byte
 * stream differs: filter choice and deflate level are not part of the image).
 *
 * Return codes / p3d_last_error() as in p3d_render.h.
 */
#ifndef P3D_IMAGEIO_H_
#define P3D_IMAGEIO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Bytes of the filtered scan-line stream of one h x w image with `channels` 8-bit channels: h * (1 + w * channels). */
size_t p3d_png_scanline_bytes(int32_t h, int32_t w, int32_t channels);

/* fp32 images (N,C,H,W), C in {1,3,4} -> N filtered scan-line streams (device memory, N * p3d_png_scanline_bytes).
   value = clamp((x + shift[c]) * scale[c], 0, 1) * 255 truncated to uint8 - `I(t).pil()` for a float tensor
   (twodee_v1.py:184) with an optional per-channel affine in front (scale/shift: HOST arrays of c floats, read before the launch; NULL = identity), which is how the
   `xyza` image `cat([(xyz + bw/2)/bw, weights])` (generate.py:141-144) is produced without materialising it:
   pass `image` = xyz planes, `extra` = the weights plane.
     extra : optional (N,1,H,W) plane appended as the LAST channel (then C counts it: C-1 planes come from `image`). */
int p3d_image_to_png_scanlines(const float* image, const float* extra, int32_t n, int32_t c, int32_t h, int32_t w,
                               const float* scale, const float* shift, uint8_t* scanlines, void* stream);

/* Same quantisation, plain interleaved pixels (N,H,W,C) uint8 (no filter bytes) for consumers that want the array. */
int p3d_image_to_u8(const float* image, const float* extra, int32_t n, int32_t c, int32_t h, int32_t w,
                    const float* scale, const float* shift, uint8_t* out_nhwc, void* stream);

/* Synchronous host-side encoder: filtered scan-lines (host memory) -> a complete PNG file image in `out`.
   Returns P3D_EWORKSPACE (and sets *out_len to the needed size) when `cap` is too small; cap >= p3d_png_encode_bound
   always suffices.  No CUDA involved: usable (and tested) without a GPU. */
size_t p3d_png_encode_bound(int32_t h, int32_t w, int32_t channels);
int p3d_png_encode_host(const uint8_t* scanlines, int32_t h, int32_t w, int32_t channels, int32_t level,
                        uint8_t* out, size_t cap, size_t* out_len);

/* Asynchronous writer: a pool of `n_threads` host threads that deflate and write PNG files.
   compress_level: zlib 0..9 (PIL's default is 6; 1-3 keeps a 512^2 RGB image under ~5 ms of one core). */
int p3d_png_writer_create(int32_t n_threads, int32_t compress_level, void** writer);
/* Enqueue one image whose filtered scan-lines live in DEVICE memory: the D2H copy (into pinned staging owned by the
   writer) is enqueued on `stream` and the call returns at once; a worker waits for the copy, encodes and writes
   `path` (written as path + ".tmp" and renamed, so a reader never sees a partial file). */
int p3d_png_writer_submit(void* writer, const uint8_t* dev_scanlines, int32_t h, int32_t w, int32_t channels,
                          const char* path, void* stream);
/* Same with scan-lines already in HOST memory (copied before the call returns).  No CUDA involved. */
int p3d_png_writer_submit_host(void* writer, const uint8_t* host_scanlines, int32_t h, int32_t w, int32_t channels,
                               const char* path);
/* Wait until every submitted image is on disk; *n_failed = number of images that could not be written since the
   last flush (p3d_last_error() has the first failure's message). */
int p3d_png_writer_flush(void* writer, int32_t* n_failed);
int p3d_png_writer_destroy(void* writer);

#ifdef __cplusplus
}
#endif
#endif /* P3D_IMAGEIO_H_ */
