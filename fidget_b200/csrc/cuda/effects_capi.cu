// C ABI of the fidget-raster post-processing effects (kernels in effects.cu).
#include "capi_internal.h"

// ---- post-processing effects (fidget-raster/src/effects.rs) ---------------------------------------
namespace {

// Device view of an input image: the pointer itself, or a staged copy of a host image.
int32_t fx_input(fc_ctx* c, DevBuf& buf, const void* p, size_t bytes, const void** dev) {
    if (is_device_ptr(p)) { *dev = p; return FC_OK; }
    CU(buf.ensure(std::max<size_t>(bytes, 16)));
    CU(cudaMemcpyAsync(buf.p, p, bytes, cudaMemcpyHostToDevice, c->stream));
    *dev = buf.p;
    return FC_OK;
}
int32_t fx_output(fc_ctx* c, void* p, size_t bytes, void** dev) {
    if (is_device_ptr(p)) { *dev = p; return FC_OK; }
    CU(c->fx_out.ensure(std::max<size_t>(bytes, 16)));
    *dev = c->fx_out.p;
    return FC_OK;
}
int32_t fx_finish(fc_ctx* c, void* p, const void* dev, size_t bytes) {
    CU(cudaGetLastError());
    if (p != dev) CU(cudaMemcpyAsync(p, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return FC_OK;
}
int32_t fx_tables(fc_ctx* c, const float* kernel, uint32_t nk, const float* noise, uint32_t nn, const float** dk,
                  const float** dn) {
    if (!kernel || !noise || !nk || !nn) return fail(FC_ERR_INVALID, "SSAO needs a kernel and a noise table");
    if (nk > 4096) return fail(FC_ERR_INVALID, "SSAO kernel table too large (max 4096 samples)");
    CU(c->fx_tables.ensure((size_t(nk) * 3 + size_t(nn) * 2) * 4));
    float* t = c->fx_tables.as<float>();
    const cudaMemcpyKind any = cudaMemcpyDefault;
    CU(cudaMemcpyAsync(t, kernel, size_t(nk) * 12, any, c->stream));
    CU(cudaMemcpyAsync(t + size_t(nk) * 3, noise, size_t(nn) * 8, any, c->stream));
    *dk = t;
    *dn = t + size_t(nk) * 3;
    return FC_OK;
}
#define FX(call) do { if (int32_t rc_ = (call)) return rc_; } while (0)

int32_t fx_to_rgba(fc_ctx* c, int mode, const float* image, uint32_t w, uint32_t h, uint8_t* out) {
    if (!c || !image || !out) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 4, &din));
    FX(fx_output(c, out, n * 4, &dout));
    launch_to_rgba(mode, static_cast<const float*>(din), n, static_cast<uint8_t*>(dout), c->stream);
    return fx_finish(c, out, dout, n * 4);
}

}  // namespace

extern "C" {

int32_t fc_denoise_normals(fc_ctx* c, const fc_geometry_pixel* image, uint32_t w, uint32_t h, fc_geometry_pixel* out) {
    if (!c || !image || !out) return fail(FC_ERR_INVALID, "null argument");
    if (image == out) return fail(FC_ERR_INVALID, "fc_denoise_normals cannot run in place");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 16, &din));
    FX(fx_output(c, out, n * 16, &dout));
    launch_denoise_normals(static_cast<const GeoPixel*>(din), w, h, static_cast<GeoPixel*>(dout), c->stream);
    return fx_finish(c, out, dout, n * 16);
}

int32_t fc_compute_ssao(fc_ctx* c, const fc_geometry_pixel* image, uint32_t w, uint32_t h, uint32_t d,
                        const float* kernel, uint32_t n_kernel, const float* noise, uint32_t n_noise, float* out) {
    if (!c || !image || !out) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const float *dk, *dn;
    FX(fx_tables(c, kernel, n_kernel, noise, n_noise, &dk, &dn));
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 16, &din));
    FX(fx_output(c, out, n * 4, &dout));
    launch_compute_ssao(static_cast<const GeoPixel*>(din), w, h, d, dk, n_kernel, dn, n_noise,
                        static_cast<float*>(dout), c->stream);
    return fx_finish(c, out, dout, n * 4);
}

int32_t fc_blur_ssao(fc_ctx* c, const float* ssao, uint32_t w, uint32_t h, float* out) {
    if (!c || !ssao || !out) return fail(FC_ERR_INVALID, "null argument");
    if (ssao == out) return fail(FC_ERR_INVALID, "fc_blur_ssao cannot run in place");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, ssao, n * 4, &din));
    FX(fx_output(c, out, n * 4, &dout));
    launch_blur_ssao(static_cast<const float*>(din), w, h, static_cast<float*>(dout), c->stream);
    return fx_finish(c, out, dout, n * 4);
}

int32_t fc_apply_shading(fc_ctx* c, const fc_geometry_pixel* image, uint32_t w, uint32_t h, uint32_t d, int32_t ssao,
                         const float* kernel, uint32_t n_kernel, const float* noise, uint32_t n_noise,
                         uint8_t* out_rgb) {
    if (!c || !image || !out_rgb) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 16, &din));
    FX(fx_output(c, out_rgb, n * 3, &dout));
    const float* occl = nullptr;
    if (ssao) {
        const float *dk, *dn;
        FX(fx_tables(c, kernel, n_kernel, noise, n_noise, &dk, &dn));
        CU(c->fx_tmp.ensure(n * 4));
        launch_compute_ssao(static_cast<const GeoPixel*>(din), w, h, d, dk, n_kernel, dn, n_noise,
                            c->fx_tmp.as<float>(), c->stream);
        occl = c->fx_tmp.as<float>();
    }
    // the 3x3-window blur of the occlusion map is applied inside the shading kernel
    launch_apply_shading(static_cast<const GeoPixel*>(din), w, h, d, occl, 1, static_cast<uint8_t*>(dout), c->stream);
    return fx_finish(c, out_rgb, dout, n * 3);
}

int32_t fc_shade_with_occlusion(fc_ctx* c, const fc_geometry_pixel* image, uint32_t w, uint32_t h, uint32_t d,
                                const float* blurred_ssao, uint8_t* out_rgb) {
    if (!c || !image || !out_rgb) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void *din, *docc = nullptr; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 16, &din));
    if (blurred_ssao) FX(fx_input(c, c->fx_tmp, blurred_ssao, n * 4, &docc));
    FX(fx_output(c, out_rgb, n * 3, &dout));
    launch_apply_shading(static_cast<const GeoPixel*>(din), w, h, d, static_cast<const float*>(docc), 0,
                         static_cast<uint8_t*>(dout), c->stream);
    return fx_finish(c, out_rgb, dout, n * 3);
}

int32_t fc_normals_to_color(fc_ctx* c, const fc_geometry_pixel* image, uint32_t w, uint32_t h, uint8_t* out_rgb) {
    if (!c || !image || !out_rgb) return fail(FC_ERR_INVALID, "null argument");
    std::lock_guard<std::mutex> guard(c->mu);
    CU(cudaSetDevice(c->device));
    const size_t n = size_t(w) * h;
    if (!n) return FC_OK;
    const void* din; void* dout;
    FX(fx_input(c, c->fx_in, image, n * 16, &din));
    FX(fx_output(c, out_rgb, n * 3, &dout));
    launch_normals_to_color(static_cast<const GeoPixel*>(din), n, static_cast<uint8_t*>(dout), c->stream);
    return fx_finish(c, out_rgb, dout, n * 3);
}

int32_t fc_to_rgba_bitmap(fc_ctx* c, const float* image, uint32_t w, uint32_t h, int32_t transparent, uint8_t* out) {
    return fx_to_rgba(c, transparent ? 1 : 0, image, w, h, out);
}
int32_t fc_to_debug_bitmap(fc_ctx* c, const float* image, uint32_t w, uint32_t h, uint8_t* out) {
    return fx_to_rgba(c, 2, image, w, h, out);
}
int32_t fc_to_rgba_distance(fc_ctx* c, const float* image, uint32_t w, uint32_t h, uint8_t* out) {
    return fx_to_rgba(c, 3, image, w, h, out);
}

}  // extern "C"
