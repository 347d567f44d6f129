// RPN training targets on the host: the reference computes these on the CPU as well
// (Cython lib/utils/bbox.pyx + numpy lib/rpn_msr/anchor_target_layer_tf.py), once per training image.
//
//   ctpn_bbox_overlaps_host / ctpn_bbox_intersections_host   the two Cython entry points, float64, same operation order
//   ctpn_anchor_targets_host   anchor_target_layer_tf.py:78-175 and :201 in one pass over the anchor grid: the anchors
//       are generated on the fly, the [inside anchors x ground truth] overlap matrix is never materialised (a CTPN anchor
//       is one 16-px column, so each feature column only visits the ground-truth strips that touch it), and labels and
//       regression targets come out for all H*W*A anchors.  Sub-sampling (numpy's global RNG) and the weight tensors stay
//       in the Python mirror so that the random stream is the reference's.
//
// Plain C++ (no CUDA): compiled with -ffp-contract=off so that every float64 expression rounds as the reference's does.
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../include/ctpn_b200.h"

namespace ctpn {
void set_error(const char *fmt, ...);      // core.cu
}

namespace {

const int kAnchorHeights[10] = {11, 16, 23, 33, 48, 68, 97, 139, 198, 283};     // generate_anchors.py:26
const int kNumAnchors = 10;

#define TT_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      ::ctpn::set_error(__VA_ARGS__);  \
      return CTPN_ERR_INVALID;         \
    }                                  \
  } while (0)

struct Box {
  double x1, y1, x2, y2;
};

inline double dmin(double a, double b) { return a < b ? a : b; }
inline double dmax(double a, double b) { return a > b ? a : b; }
inline double area(const Box &b) { return (b.x2 - b.x1 + 1) * (b.y2 - b.y1 + 1); }

// bbox.pyx:38-54 (n = `boxes` row, q = `query_boxes` row); 0 where the boxes do not intersect
inline double overlap(const Box &n, const Box &q, double q_area) {
  const double iw = dmin(n.x2, q.x2) - dmax(n.x1, q.x1) + 1;
  if (!(iw > 0)) return 0.0;
  const double ih = dmin(n.y2, q.y2) - dmax(n.y1, q.y1) + 1;
  if (!(ih > 0)) return 0.0;
  const double ua = area(n) + q_area - iw * ih;
  return iw * ih / ua;
}

// bbox.pyx:80-94
inline double intersection(const Box &n, const Box &q, double q_area) {
  const double iw = dmin(n.x2, q.x2) - dmax(n.x1, q.x1) + 1;
  if (!(iw > 0)) return 0.0;
  const double ih = dmin(n.y2, q.y2) - dmax(n.y1, q.y1) + 1;
  if (!(ih > 0)) return 0.0;
  return iw * ih / q_area;
}

inline Box load(const double *p) { return {p[0], p[1], p[2], p[3]}; }

// generate_anchors.py:13-21 stores float expressions into an int32 array (truncation toward zero)
inline Box anchor_at(int row, int col, int a, int stride) {
  const int h = kAnchorHeights[a];
  const double y_lo = 7.5 - h / 2.0, y_hi = 7.5 + h / 2.0;
  return {(double)(col * stride + 0), (double)(row * stride + (int)y_lo), (double)(col * stride + 15),
          (double)(row * stride + (int)y_hi)};
}

template <bool kIntersections>
int pairwise(const double *boxes, int n, int boxes_stride, const double *query, int k, int query_stride, double *out) {
  TT_REQUIRE(n >= 0 && k >= 0 && boxes_stride >= 4 && query_stride >= 4, "bbox pairwise: bad arguments");
  TT_REQUIRE((n == 0 || boxes) && (k == 0 || query) && (n == 0 || k == 0 || out), "bbox pairwise: null pointer");
  for (int j = 0; j < k; ++j) {
    const Box q = load(query + (size_t)j * query_stride);
    const double q_area = area(q);
    for (int i = 0; i < n; ++i) {
      const Box b = load(boxes + (size_t)i * boxes_stride);
      out[(size_t)i * k + j] = kIntersections ? intersection(b, q, q_area) : overlap(b, q, q_area);
    }
  }
  return CTPN_OK;
}

}  // namespace

extern "C" int ctpn_bbox_overlaps_host(const double *boxes, int n, int boxes_stride, const double *query_boxes, int k,
                                       int query_stride, double *overlaps) {
  return pairwise<false>(boxes, n, boxes_stride, query_boxes, k, query_stride, overlaps);
}

extern "C" int ctpn_bbox_intersections_host(const double *boxes, int n, int boxes_stride, const double *query_boxes, int k,
                                            int query_stride, double *intersections) {
  return pairwise<true>(boxes, n, boxes_stride, query_boxes, k, query_stride, intersections);
}

extern "C" int ctpn_anchor_targets_host(const double *gt_boxes, int num_gt, int gt_is_f32, const unsigned char *gt_ishard,
                                        const double *dontcare_areas, int num_dontcare, int feat_h, int feat_w,
                                        int feat_stride, double im_h, double im_w, const double *cfg5, float *labels,
                                        float *bbox_targets) {
  TT_REQUIRE(gt_boxes && cfg5 && labels && bbox_targets, "ctpn_anchor_targets_host: null pointer");
  TT_REQUIRE(num_gt > 0, "ctpn_anchor_targets_host: no ground-truth boxes (the reference's argmax over an empty axis raises too)");
  TT_REQUIRE(feat_h > 0 && feat_w > 0 && feat_stride > 0 && num_dontcare >= 0 && (num_dontcare == 0 || dontcare_areas),
             "ctpn_anchor_targets_host: bad arguments");
  const double neg_thr = cfg5[0], pos_thr = cfg5[1], dontcare_hi = cfg5[3];
  const bool clobber = cfg5[2] != 0, preclude_hard = cfg5[4] != 0;
  const size_t total = (size_t)feat_h * feat_w * kNumAnchors;
  for (size_t i = 0; i < total; ++i) labels[i] = -1.0f;
  for (size_t i = 0; i < 4 * total; ++i) bbox_targets[i] = 0.0f;

  std::vector<Box> gt(num_gt);
  std::vector<double> gt_area(num_gt);
  for (int k = 0; k < num_gt; ++k) {
    gt[k] = load(gt_boxes + 4 * (size_t)k);
    gt_area[k] = area(gt[k]);
  }
  // ground-truth boxes that touch each feature column in x (iw > 0), ascending index so that ties resolve as argmax does
  std::vector<std::vector<int>> touching(feat_w);
  for (int col = 0; col < feat_w; ++col) {
    const double ax1 = (double)col * feat_stride, ax2 = ax1 + 15;
    for (int k = 0; k < num_gt; ++k)
      if (dmin(ax2, gt[k].x2) - dmax(ax1, gt[k].x1) + 1 > 0) touching[col].push_back(k);
  }
  // anchor_target_layer_tf.py:101-106: only anchors inside the image take part (allowed border 0)
  auto inside = [&](const Box &a) { return a.x1 >= 0 && a.y1 >= 0 && a.x2 < im_w && a.y2 < im_h; };

  // pass 1: every inside anchor's best ground truth (first maximum) and every ground truth's best overlap (:127-133)
  std::vector<double> best(total, 0.0);
  std::vector<int> best_gt(total, 0);
  std::vector<double> gt_best(num_gt, 0.0);
  size_t num_inside = 0, first_inside = total;
  for (int row = 0; row < feat_h; ++row)
    for (int col = 0; col < feat_w; ++col)
      for (int a = 0; a < kNumAnchors; ++a) {
        const Box an = anchor_at(row, col, a, feat_stride);
        if (!inside(an)) continue;
        const size_t i = ((size_t)row * feat_w + col) * kNumAnchors + a;
        if (first_inside == total) first_inside = i;
        ++num_inside;
        double b = 0.0;
        int arg = 0;
        for (int k : touching[col]) {
          const double ov = overlap(an, gt[k], gt_area[k]);
          if (ov > b) b = ov, arg = k;
          if (ov > gt_best[k]) gt_best[k] = ov;
        }
        best[i] = b;
        best_gt[i] = arg;
      }
  TT_REQUIRE(num_inside > 0, "ctpn_anchor_targets_host: no anchor lies inside the %g x %g image", im_w, im_h);
  // :134-136 `overlaps == gt_max_overlaps`: a ground truth nothing overlaps has maximum 0 and ties with every anchor
  bool unmatched_gt = false;
  for (int k = 0; k < num_gt; ++k) unmatched_gt |= (gt_best[k] == 0.0);

  std::vector<float> log_cache((size_t)num_gt * (kNumAnchors + 1));
  std::vector<unsigned char> log_have((size_t)num_gt * (kNumAnchors + 1), 0);
  // pass 2: labels (:138-150) and regression targets against the best ground truth (:201, bbox_transform.py:10-29)
  for (int row = 0; row < feat_h; ++row)
    for (int col = 0; col < feat_w; ++col)
      for (int a = 0; a < kNumAnchors; ++a) {
        const Box an = anchor_at(row, col, a, feat_stride);
        if (!inside(an)) continue;
        const size_t i = ((size_t)row * feat_w + col) * kNumAnchors + a;
        bool tied = unmatched_gt;
        if (!tied)
          for (int k : touching[col])
            if (overlap(an, gt[k], gt_area[k]) == gt_best[k]) {
              tied = true;
              break;
            }
        float lab = -1.0f;
        if (!clobber && best[i] < neg_thr) lab = 0.0f;
        if (tied) lab = 1.0f;
        if (best[i] >= pos_thr) lab = 1.0f;
        if (clobber && best[i] < neg_thr) lab = 0.0f;
        labels[i] = lab;

        const Box &g = gt[best_gt[i]];
        const double ex_w = an.x2 - an.x1 + 1.0, ex_h = an.y2 - an.y1 + 1.0;
        const double ex_cx = an.x1 + 0.5 * ex_w, ex_cy = an.y1 + 0.5 * ex_h;
        double gt_w, gt_h, gt_cx, gt_cy;
        if (gt_is_f32) {     // float32 annotations: numpy keeps the ground-truth side of bbox_transform in float32
          const float x1 = (float)g.x1, y1 = (float)g.y1, x2 = (float)g.x2, y2 = (float)g.y2;
          const float w = x2 - x1 + 1.0f, h = y2 - y1 + 1.0f;
          gt_w = w, gt_h = h, gt_cx = x1 + 0.5f * w, gt_cy = y1 + 0.5f * h;
        } else {
          gt_w = g.x2 - g.x1 + 1.0, gt_h = g.y2 - g.y1 + 1.0;
          gt_cx = g.x1 + 0.5 * gt_w, gt_cy = g.y1 + 0.5 * gt_h;
        }
        float *t = bbox_targets + 4 * i;
        t[0] = (float)((gt_cx - ex_cx) / ex_w);
        t[1] = (float)((gt_cy - ex_cy) / ex_h);
        // the two logarithms depend on (ground truth, anchor shape) only -- every anchor is 16 wide and has one of 10
        // heights -- so each distinct value is computed once instead of twice per anchor
        float *cached = &log_cache[(size_t)best_gt[i] * (kNumAnchors + 1)];
        unsigned char *have = &log_have[(size_t)best_gt[i] * (kNumAnchors + 1)];
        if (!have[kNumAnchors]) cached[kNumAnchors] = (float)log(gt_w / ex_w), have[kNumAnchors] = 1;
        if (!have[a]) cached[a] = (float)log(gt_h / ex_h), have[a] = 1;
        t[2] = cached[kNumAnchors];
        t[3] = cached[a];
      }

  // :153-160 dontcare areas: share of each anchor covered by them, summed area by area
  if (num_dontcare > 0) {
    for (int row = 0; row < feat_h; ++row)
      for (int col = 0; col < feat_w; ++col)
        for (int a = 0; a < kNumAnchors; ++a) {
          const Box an = anchor_at(row, col, a, feat_stride);
          if (!inside(an)) continue;
          const double an_area = area(an);
          double cover = 0.0;
          for (int d = 0; d < num_dontcare; ++d) cover += intersection(load(dontcare_areas + 4 * (size_t)d), an, an_area);
          if (cover > dontcare_hi) labels[((size_t)row * feat_w + col) * kNumAnchors + a] = -1.0f;
        }
  }

  // :164-177 hard ground truth: anchors that match it well, and its best anchor (first maximum; anchor 0 of the inside
  // set when nothing overlaps it), are ignored
  if (preclude_hard && gt_ishard) {
    for (int k = 0; k < num_gt; ++k) {
      if (gt_ishard[k] != 1) continue;
      double hb = 0.0;
      size_t harg = first_inside;
      for (int row = 0; row < feat_h; ++row)
        for (int col = 0; col < feat_w; ++col) {
          const double ax1 = (double)col * feat_stride, ax2 = ax1 + 15;
          if (!(dmin(ax2, gt[k].x2) - dmax(ax1, gt[k].x1) + 1 > 0)) continue;
          for (int a = 0; a < kNumAnchors; ++a) {
            const Box an = anchor_at(row, col, a, feat_stride);
            if (!inside(an)) continue;
            const size_t i = ((size_t)row * feat_w + col) * kNumAnchors + a;
            const double ov = overlap(gt[k], an, area(an));
            if (ov >= pos_thr) labels[i] = -1.0f;
            if (ov > hb) hb = ov, harg = i;
          }
        }
      labels[harg] = -1.0f;
    }
  }
  return CTPN_OK;
}
