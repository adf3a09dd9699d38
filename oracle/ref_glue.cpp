// TEST INFRASTRUCTURE ONLY: C entry point around the REFERENCE's own boxes_iou_bev_cpu (pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:232-252),
// compiled from where it lies under /root/reference by oracle/build_ref.py.  No reference source is copied into this repository.
#include <torch/torch.h>
int boxes_iou_bev_cpu(at::Tensor boxes_a_tensor, at::Tensor boxes_b_tensor, at::Tensor ans_iou_tensor);
extern "C" int ref_boxes_iou_bev(const float* a, int64_t na, const float* b, int64_t nb, float* out) {
  auto opt = torch::TensorOptions().dtype(torch::kFloat32);
  at::Tensor ta = torch::from_blob((void*)a, {na, 7}, opt), tb = torch::from_blob((void*)b, {nb, 7}, opt), to = torch::from_blob((void*)out, {na, nb}, opt);
  return boxes_iou_bev_cpu(ta, tb, to);
}
