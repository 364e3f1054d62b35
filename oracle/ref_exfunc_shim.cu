// extern "C" re-export of the REFERENCE's own point_render / point_masker CUDA launchers (TEST INFRASTRUCTURE).
// The reference sources are compiled where they lie (/root/reference/model/exfunc/expackages/*/*_cuda.cu, see
// oracle/Makefile, target _ref/libexfunc_ref.so); nothing of them is copied into this repository.  Only the raw
// kernels + launchers compile standalone (the pybind/torch wrappers *.cpp need the torch C++ API and are restated
// in tests/test_gpu_exfunc.py: allocation shapes and initial values of point_render.cpp:80-88, 119-120 and
// point_masker.cpp:55, 81).
void point_render_cuda_forward(float* proj_points, int* pixel_index, int* is_visible, float* front_image,
                               float* back_image, float* weight_points, float* depth_image, float* weight_image,
                               int batch_size, int point_num, int height, int width, int visible_patch_width,
                               int color_patch_width, float threshold);
void point_render_cuda_backward(float* grad_depth_image, float* grad_weight_image, float* proj_points,
                                int* pixel_index, int* is_visible, float* weight_points, float* grad_weight_points,
                                float* grad_proj_points, int batch_size, int point_num, int height, int width,
                                int color_patch_width);
void point_masker_cuda_forward(float* proj_points, float* mask_image, int batch_size, int point_num, int height,
                               int width, int mask_patch_width);
void point_masker_cuda_backward(float* grad_mask_image, float* proj_points, float* grad_proj_points, int batch_size,
                                int point_num, int height, int width, float epsilon, float threshold);

extern "C" {
void ref_point_render_forward(float* proj_points, int* pixel_index, int* is_visible, float* front_image,
                              float* back_image, float* weight_points, float* depth_image, float* weight_image,
                              int batch_size, int point_num, int height, int width, int visible_patch_width,
                              int color_patch_width, float threshold) {
  point_render_cuda_forward(proj_points, pixel_index, is_visible, front_image, back_image, weight_points, depth_image,
                            weight_image, batch_size, point_num, height, width, visible_patch_width,
                            color_patch_width, threshold);
}
void ref_point_render_backward(float* grad_depth_image, float* grad_weight_image, float* proj_points,
                               int* pixel_index, int* is_visible, float* weight_points, float* grad_weight_points,
                               float* grad_proj_points, int batch_size, int point_num, int height, int width,
                               int color_patch_width) {
  point_render_cuda_backward(grad_depth_image, grad_weight_image, proj_points, pixel_index, is_visible,
                             weight_points, grad_weight_points, grad_proj_points, batch_size, point_num, height,
                             width, color_patch_width);
}
void ref_point_masker_forward(float* proj_points, float* mask_image, int batch_size, int point_num, int height,
                              int width, int mask_patch_width) {
  point_masker_cuda_forward(proj_points, mask_image, batch_size, point_num, height, width, mask_patch_width);
}
void ref_point_masker_backward(float* grad_mask_image, float* proj_points, float* grad_proj_points, int batch_size,
                               int point_num, int height, int width, float epsilon, float threshold) {
  point_masker_cuda_backward(grad_mask_image, proj_points, grad_proj_points, batch_size, point_num, height, width,
                             epsilon, threshold);
}
}
