// Reference-side binding of libgsr_b200: what $RAST/rasterize_points.cu becomes when a maintainer keeps the
// reference's pybind layer (ext.cpp:15-19, rasterize_points.h:18-70) and replaces only the torch-free core
// (CudaRasterizer::Rasterizer, cuda_rasterizer/rasterizer.h:24-91) with the C ABI of include/gsr.h.
//
// Same three functions, same argument lists, same returned tuples as rasterize_points.cu:35-231; plain C++ (no
// CUDA in this translation unit).  Build: add `-I<repo>/include -L<repo>/gaustudio_b200 -lgsr_b200
// -Wl,-rpath,<repo>/gaustudio_b200` to $RAST/setup.py:29, list this file instead of rasterize_points.cu and drop
// cuda_rasterizer/*.cu.  tests/test_abi.py compiles this file against include/gsr.h and the torch headers.
#include <torch/extension.h>

#include <c10/cuda/CUDAStream.h>

#include <string>
#include <tuple>

#include "gsr.h"

namespace {

// replaces resizeFunctional (rasterize_points.cu:27-33): grow the byte tensor, hand its storage to the library
char* grow(void* user, size_t bytes) {
  auto* t = static_cast<torch::Tensor*>(user);
  t->resize_({static_cast<int64_t>(bytes)});
  return reinterpret_cast<char*>(t->data_ptr());
}

// device pointer of an optional input: the reference passes `.contiguous().data<float>()`, which is nullptr for
// the empty tensor that stands for "absent" (__init__.py:200-212)
struct F32 {
  torch::Tensor keep;
  explicit F32(const torch::Tensor& t) : keep(t.numel() ? t.contiguous() : t) {}
  const float* ptr() const { return keep.numel() ? keep.data_ptr<float>() : nullptr; }
};

void* current_stream() { return static_cast<void*>(c10::cuda::getCurrentCUDAStream().stream()); }

}  // namespace

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                       const torch::Tensor& opacity, const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                       const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                       const int image_height, const int image_width, const torch::Tensor& sh, const int degree,
                       const torch::Tensor& campos, const bool prefiltered, const bool debug) {
  if (means3D.ndimension() != 2 || means3D.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
  const int P = static_cast<int>(means3D.size(0));
  const int H = image_height, W = image_width;
  auto f32 = means3D.options().dtype(torch::kFloat32);
  torch::Tensor out_color = torch::empty({3, H, W}, f32), out_depth = torch::empty({1, H, W}, f32);
  torch::Tensor out_median = torch::empty({3, H, W}, f32), out_opacity = torch::empty({1, H, W}, f32);
  torch::Tensor radii = torch::empty({P}, means3D.options().dtype(torch::kInt32));
  auto bytes = means3D.options().dtype(torch::kByte);
  torch::Tensor geomBuffer = torch::empty({0}, bytes), binningBuffer = torch::empty({0}, bytes),
                imgBuffer = torch::empty({0}, bytes);
  int64_t rendered = 0;
  if (P != 0) {
    const int M = sh.numel() ? static_cast<int>(sh.size(1)) : 0;  // rasterize_points.cu:86-90
    // gaustudio keeps the background colour on the CPU (vanilla_renderer.py:23); the library reads it on the device
    F32 bg(background.to(means3D.device())), m3(means3D), shs(sh), col(colors), opa(opacity), sc(scales), rot(rotations),
        cov(cov3D_precomp), view(viewmatrix), proj(projmatrix), cam(campos);
    rendered = gsr_forward(grow, &geomBuffer, grow, &binningBuffer, grow, &imgBuffer, P, degree, M, bg.ptr(), W, H,
                           m3.ptr(), shs.ptr(), col.ptr(), opa.ptr(), sc.ptr(), scale_modifier, rot.ptr(), cov.ptr(),
                           view.ptr(), proj.ptr(), cam.ptr(), tan_fovx, tan_fovy, prefiltered ? 1 : 0,
                           out_color.data_ptr<float>(), out_depth.data_ptr<float>(), out_median.data_ptr<float>(),
                           out_opacity.data_ptr<float>(), radii.data_ptr<int>(), debug ? 1 : 0,
                           /*r_capacity=*/0, /*r_host=*/nullptr, current_stream());
    if (rendered < 0) AT_ERROR(std::string("gsr_forward: ") + gsr_last_error());
  } else {
    out_color.zero_(); out_depth.zero_(); out_median.zero_(); out_opacity.zero_();
  }
  return std::make_tuple(static_cast<int>(rendered), out_color, out_depth, out_median, out_opacity, radii, geomBuffer,
                         binningBuffer, imgBuffer);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                               const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                               const float scale_modifier, const torch::Tensor& cov3D_precomp,
                               const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
                               const float tan_fovy, const torch::Tensor& dL_dout_color,
                               const torch::Tensor& dL_dout_depth, const torch::Tensor& dL_dout_median_depth,
                               const torch::Tensor& dL_dout_final_opacity, const torch::Tensor& sh, const int degree,
                               const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer, const bool debug) {
  const int P = static_cast<int>(means3D.size(0));
  const int H = static_cast<int>(dL_dout_color.size(1)), W = static_cast<int>(dL_dout_color.size(2));
  const int M = sh.numel() ? static_cast<int>(sh.size(1)) : 0;
  auto f32 = means3D.options().dtype(torch::kFloat32);
  // every element is written by the library when P > 0: no torch::zeros pre-pass (rasterize_points.cu:160-169)
  auto make = [&](std::initializer_list<int64_t> shape) { return P ? torch::empty(shape, f32) : torch::zeros(shape, f32); };
  torch::Tensor dL_dmeans3D = make({P, 3}), dL_dmeans2D = make({P, 3}), dL_dcolors = make({P, 3});
  torch::Tensor dL_dopacity = make({P, 1}), dL_dcov3D = make({P, 6}), dL_dsh = make({P, M, 3});
  torch::Tensor dL_dscales = make({P, 3}), dL_drotations = make({P, 4});
  if (P != 0) {
    F32 bg(background.to(means3D.device())), m3(means3D), shs(sh), col(colors), sc(scales), rot(rotations),
        cov(cov3D_precomp), view(viewmatrix), proj(projmatrix), cam(campos), gc(dL_dout_color), gd(dL_dout_depth),
        gm(dL_dout_median_depth), go(dL_dout_final_opacity);
    torch::Tensor rad = radii.contiguous();
    const int rc = gsr_backward(
        P, degree, M, static_cast<int64_t>(R), bg.ptr(), W, H, m3.ptr(), shs.ptr(), col.ptr(), sc.ptr(), scale_modifier,
        rot.ptr(), cov.ptr(), view.ptr(), proj.ptr(), cam.ptr(), tan_fovx, tan_fovy, rad.data_ptr<int>(),
        reinterpret_cast<char*>(geomBuffer.data_ptr()), reinterpret_cast<char*>(binningBuffer.data_ptr()),
        reinterpret_cast<char*>(imageBuffer.data_ptr()), gc.ptr(), gd.ptr(), gm.ptr(), go.ptr(),
        dL_dmeans2D.data_ptr<float>(), /*dL_dconic=*/nullptr, dL_dopacity.data_ptr<float>(), dL_dcolors.data_ptr<float>(),
        /*dL_ddepth=*/nullptr, dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(),
        M ? dL_dsh.data_ptr<float>() : nullptr, dL_dscales.data_ptr<float>(), dL_drotations.data_ptr<float>(),
        debug ? 1 : 0, current_stream());
    if (rc < 0) AT_ERROR(std::string("gsr_backward: ") + gsr_last_error());
  }
  return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations);
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix) {
  const int P = static_cast<int>(means3D.size(0));
  torch::Tensor present = torch::zeros({P}, means3D.options().dtype(torch::kBool));
  if (P != 0) {
    F32 m3(means3D), view(viewmatrix), proj(projmatrix);
    const int rc = gsr_mark_visible(P, m3.ptr(), view.ptr(), proj.ptr(),
                                    reinterpret_cast<unsigned char*>(present.data_ptr<bool>()), current_stream());
    if (rc < 0) AT_ERROR(std::string("gsr_mark_visible: ") + gsr_last_error());
  }
  return present;
}

// the reference's module definition, unchanged (ext.cpp:15-19)
PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("rasterize_gaussians", &RasterizeGaussiansCUDA);
  m.def("rasterize_gaussians_backward", &RasterizeGaussiansBackwardCUDA);
  m.def("mark_visible", &markVisible);
}
