// extern "C" doorway to the REFERENCE's own `_nms` (lib/utils/nms_kernel.cu:91-144), which is a
// C++-mangled symbol (lib/utils/gpu_nms.hpp:1-2).  Compiled together with the reference source
// where it lies under /root/reference by oracle/Makefile; nothing of the reference is copied.
void _nms(int *keep_out, int *num_out, const float *boxes_host, int boxes_num, int boxes_dim,
          float nms_overlap_thresh, int device_id);

extern "C" void ref_nms(int *keep_out, int *num_out, const float *boxes_host, int boxes_num, int boxes_dim,
                        float nms_overlap_thresh, int device_id) {
  _nms(keep_out, num_out, boxes_host, boxes_num, boxes_dim, nms_overlap_thresh, device_id);
}
