// segment.cpp -- see segment.hpp.  Every function cites the reference lines (or the OpenCV
// call made there) whose results it reproduces.
#include "segment.hpp"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <mutex>
#include <random>
#include <string>
#include <thread>

#include "../../include/svc_segment.h"

static_assert(sizeof(Vec2f) == 2 * sizeof(float), "Vec2f must be two packed floats");

namespace svc {

// ---- generators ----------------------------------------------------------------------
// libstdc++ std::uniform_int_distribution<unsigned>{0, hi}(std::minstd_rand0&): the generator's
// range is [1, 2147483646] (not a power of two), so the library scales it down with two
// divisions and rejects the tail.
uint32_t MinstdRand0::UniformInclusive(uint32_t hi) {
  const uint64_t urng_range = 2147483646ull - 1ull;
  const uint64_t ue_range = (uint64_t)hi + 1ull;
  if (urng_range > hi) {
    const uint64_t scaling = urng_range / ue_range;
    const uint64_t past = ue_range * scaling;
    uint64_t ret;
    do {
      ret = (uint64_t)Next() - 1ull;
    } while (ret >= past);
    return (uint32_t)(ret / scaling);
  }
  // hi >= 2^31 - 3 never happens for motion fields; keep the call total all the same
  return (uint32_t)(((uint64_t)Next() - 1ull) % ue_range);
}

// ---- RANSAC (libs/motion.cpp:157-266) ---------------------------------------------------
static uint RansacIterCount(const RansacParams& p) {  // libs/motion.cpp:157-162
  const float quot = std::log(1 - p.success_prob);
  const float div = std::log(1 - std::pow(p.inlier_ratio, (float)p.subset_sz));
  return (uint)std::ceil(quot / div);
}

static Vec2f AvgOf(const Vec2f* mf, const uint* idx, uint n) {  // libs/motion.cpp:164-177
  Vec2f r{0.f, 0.f};
  for (uint i = 0; i < n; ++i) {
    r.x += mf[idx[i]].x;
    r.y += mf[idx[i]].y;
  }
  const float s = 1.0f / n;
  r.x = r.x * s;
  r.y = r.y * s;
  return r;
}

static float RmseOf(const Vec2f* mf, const uint* idx, uint n, Vec2f est) {  // libs/motion.cpp:179-195
  float r = 0;
  for (uint i = 0; i < n; ++i) {
    const Vec2f m = mf[idx[i]];
    const float ex = m.x - est.x, ey = m.y - est.y;
    r += ex * ex + ey * ey;
  }
  return std::sqrt(r / n);
}

void RansacGlobalMotion(const Vec2f* mf, uint n, const RansacParams& params, MinstdRand0& rng, float* rmse,
                        Vec2f* global_motion, std::vector<uint>* inliers_out) {
  const uint iters = RansacIterCount(params);
  const float thr2 = params.inlier_thresh * params.inlier_thresh;
  std::vector<uint> subset(params.subset_sz), best_subset(params.subset_sz);
  std::vector<uint> inliers, best;
  inliers.reserve(n);
  best.reserve(n);
  Vec2f best_gm{0.f, 0.f};
  for (uint it = 0; it < iters; ++it) {
    // distinct indices, drawn from [0, n] INCLUSIVE exactly like the reference (libs/motion.cpp:208),
    // which reads one element past the field when n comes up.  That iteration cannot be evaluated:
    // it is dropped here, after consuming the same draws, so the sample stream stays the reference's.
    bool past_end = false;
    for (uint i = 0; i < params.subset_sz; ++i) {
      uint j;
      do {
        subset[i] = rng.UniformInclusive(n);
        j = 0;
        while (j < i && subset[j] != subset[i]) ++j;
      } while (j < i);
      past_end |= subset[i] == n;
    }
    if (past_end) continue;
    const Vec2f gm = AvgOf(mf, subset.data(), (uint)subset.size());
    inliers.clear();
    for (uint i = 0; i < n; ++i) {
      const float dx = gm.x - mf[i].x, dy = gm.y - mf[i].y;
      if (dx * dx + dy * dy < thr2) inliers.push_back(i);
    }
    if (inliers.size() >= best.size()) {  // ">=": a later subset with as many inliers wins
      best_gm = gm;
      best_subset.swap(subset);
      best.swap(inliers);
    }
  }
  if (best.size() < params.subset_sz) {
    // degenerate path of the reference (libs/motion.cpp:239-241): error against the CALLER's value
    *rmse = RmseOf(mf, best_subset.data(), params.subset_sz, *global_motion);
  } else {
    // refit on the consensus set; membership is not re-evaluated against the refit (libs/motion.cpp:243-261)
    best_gm = AvgOf(mf, best.data(), (uint)best.size());
    *rmse = RmseOf(mf, best.data(), (uint)best.size(), best_gm);
  }
  *global_motion = best_gm;
  inliers_out->swap(best);
}

// ---- morphology (cv::morphologyEx with a MORPH_RECT element, libs/encoder.cpp:189-190, 519-522) ----
// OpenCV: dst(x,y) = min|max over the element's cells (i,j) of src(x + i - anchor.x, y + j - anchor.y),
// anchor = element centre (w/2, h/2); the default border value makes out-of-image cells neutral.
static void RectMinMax(const uint8_t* src, uint8_t* dst, uint w, uint h, uint rw, uint rh, bool take_max,
                       std::vector<uint8_t>& tmp) {
  const int ax = (int)rw / 2, ay = (int)rh / 2;
  tmp.resize((size_t)w * h);
  for (uint y = 0; y < h; ++y)
    for (uint x = 0; x < w; ++x) {
      const int x0 = std::max(0, (int)x - ax), x1 = std::min((int)w - 1, (int)x - ax + (int)rw - 1);
      uint8_t v = take_max ? 0 : 255;
      for (int i = x0; i <= x1; ++i) v = take_max ? std::max(v, src[y * w + i]) : std::min(v, src[y * w + i]);
      tmp[y * w + x] = v;
    }
  for (uint y = 0; y < h; ++y) {
    const int y0 = std::max(0, (int)y - ay), y1 = std::min((int)h - 1, (int)y - ay + (int)rh - 1);
    for (uint x = 0; x < w; ++x) {
      uint8_t v = take_max ? 0 : 255;
      for (int j = y0; j <= y1; ++j) v = take_max ? std::max(v, tmp[j * w + x]) : std::min(v, tmp[j * w + x]);
      dst[y * w + x] = v;
    }
  }
}

void MorphologyEx(uint8_t* mask, uint w, uint h, uint op, uint rw, uint rh) {
  std::vector<uint8_t> tmp;
  switch (op) {
    case kErode: RectMinMax(mask, mask, w, h, rw, rh, false, tmp); break;
    case kDilate: RectMinMax(mask, mask, w, h, rw, rh, true, tmp); break;
    case kOpen:
      RectMinMax(mask, mask, w, h, rw, rh, false, tmp);
      RectMinMax(mask, mask, w, h, rw, rh, true, tmp);
      break;
    default:  // kClose
      RectMinMax(mask, mask, w, h, rw, rh, true, tmp);
      RectMinMax(mask, mask, w, h, rw, rh, false, tmp);
      break;
  }
}

// ---- connected components (cv::connectedComponents, libs/encoder.cpp:607-611) -------------------
// Two-pass union-find.  OpenCV numbers the components by the order in which its scan creates
// their first provisional label: pixel raster order for 4-connectivity (SAUF / Spaghetti4C),
// raster order of 2x2 pixel blocks for 8-connectivity (BBDT / Spaghetti scan 2x2 blocks).
static int32_t Find(std::vector<int32_t>& parent, int32_t a) {
  while (parent[a] != a) {
    parent[a] = parent[parent[a]];
    a = parent[a];
  }
  return a;
}

uint ConnectedComponents(const uint8_t* mask, uint w, uint h, uint connectivity, int32_t* labels) {
  std::vector<int32_t> parent(1, 0);
  auto unite = [&](int32_t a, int32_t b) {
    a = Find(parent, a);
    b = Find(parent, b);
    if (a != b) parent[std::max(a, b)] = std::min(a, b);
    return std::min(a, b);
  };
  for (uint y = 0; y < h; ++y)
    for (uint x = 0; x < w; ++x) {
      int32_t& out = labels[y * w + x];
      out = 0;
      if (!mask[y * w + x]) continue;
      int32_t l = 0;
      auto look = [&](int xx, int yy) {
        if (xx < 0 || yy < 0 || xx >= (int)w) return;
        const int32_t o = labels[yy * (int)w + xx];
        if (o) l = l ? unite(l, o) : o;
      };
      look((int)x - 1, (int)y);
      look((int)x, (int)y - 1);
      if (connectivity == 8) {
        look((int)x - 1, (int)y - 1);
        look((int)x + 1, (int)y - 1);
      }
      if (!l) {
        l = (int32_t)parent.size();
        parent.push_back(l);
      }
      out = l;
    }
  // final numbering
  const uint bw = (w + 1) / 2;
  std::vector<uint64_t> first(parent.size(), ~0ull);  // per root: smallest scan key of its pixels
  for (uint y = 0; y < h; ++y)
    for (uint x = 0; x < w; ++x) {
      int32_t& l = labels[y * w + x];
      if (!l) continue;
      l = Find(parent, l);
      const uint64_t key = connectivity == 8 ? (uint64_t)(y / 2) * bw + x / 2 : (uint64_t)y * w + x;
      first[l] = std::min(first[l], key);
    }
  std::vector<int32_t> roots;
  for (size_t i = 1; i < parent.size(); ++i)
    if (parent[i] == (int32_t)i && first[i] != ~0ull) roots.push_back((int32_t)i);
  std::sort(roots.begin(), roots.end(), [&](int32_t a, int32_t b) { return first[a] < first[b]; });
  std::vector<int32_t> final_id(parent.size(), 0);
  for (size_t i = 0; i < roots.size(); ++i) final_id[roots[i]] = (int32_t)i + 1;
  for (size_t i = 0; i < (size_t)w * h; ++i)
    if (labels[i]) labels[i] = final_id[labels[i]];
  return (uint)roots.size() + 1;
}

// ---- k-means (cv::kmeans, KMEANS_PP_CENTERS; libs/encoder.cpp:570-575) --------------------------
static inline float NormL2Sqr(const float* a, const float* b, int n) {  // cv::hal::normL2Sqr_, scalar tail
  float d = 0.f;
  for (int j = 0; j < n; ++j) {
    const float t = a[j] - b[j];
    d += t * t;
  }
  return d;
}

// Distances of all N samples to one point / of one sample to all K centres, with the samples
// (centres) stored one coordinate per array: the loops vectorise over samples (centres) while every
// distance keeps NormL2Sqr's operation order ((t0^2 + t1^2) + t2^2) + ..., so the floats are the
// same as the scalar form's.
struct Soa {
  int n = 0, dims = 0;
  std::vector<float> v;  // dims x n
  void Load(const float* aos, int n_, int dims_) {
    n = n_;
    dims = dims_;
    v.resize((size_t)n * dims);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < dims; ++j) v[(size_t)j * n + i] = aos[(size_t)i * dims + j];
  }
  const float* Row(int j) const { return v.data() + (size_t)j * n; }
};

static void DistToPoint(const Soa& s, const float* __restrict__ pt, float* __restrict__ out) {
  const int n = s.n;
  if (s.dims == 4) {
    const float *x0 = s.Row(0), *x1 = s.Row(1), *x2 = s.Row(2), *x3 = s.Row(3);
    const float p0 = pt[0], p1 = pt[1], p2 = pt[2], p3 = pt[3];
    for (int i = 0; i < n; ++i) {
      const float t0 = x0[i] - p0, t1 = x1[i] - p1, t2 = x2[i] - p2, t3 = x3[i] - p3;
      out[i] = ((0.f + t0 * t0) + t1 * t1 + t2 * t2) + t3 * t3;
    }
    return;
  }
  for (int i = 0; i < n; ++i) out[i] = 0.f;
  for (int j = 0; j < s.dims; ++j) {
    const float* x = s.Row(j);
    const float p = pt[j];
    for (int i = 0; i < n; ++i) {
      const float t = x[i] - p;
      out[i] += t * t;
    }
  }
}

// k-means++ seeding with 3 trials per centre (Arthur & Vassilvitskii; OpenCV generateCentersPP)
static void CentersPP(const float* data, const Soa& soa, int N, int dims, int K, CvRng& rng, float* out_centers,
                      std::vector<float>& buf, std::vector<int>& chosen) {
  const int trials = 3;
  buf.resize((size_t)N * 4);
  chosen.resize(K);
  float *dist = buf.data(), *tdist = dist + N, *tdist2 = tdist + N, *raw = tdist2 + N;
  double sum0 = 0;
  chosen[0] = (int)(rng.Next() % (uint32_t)N);
  DistToPoint(soa, data + (size_t)chosen[0] * dims, dist);
  for (int i = 0; i < N; ++i) sum0 += dist[i];
  for (int k = 1; k < K; ++k) {
    double best_sum = DBL_MAX;
    int best_center = -1;
    for (int j = 0; j < trials; ++j) {
      double p = rng.NextDouble() * sum0;
      int ci = 0;
      for (; ci < N - 1; ++ci) {
        p -= dist[ci];
        if (p <= 0) break;
      }
      double s = 0;
      DistToPoint(soa, data + (size_t)ci * dims, raw);
      for (int i = 0; i < N; ++i) tdist2[i] = std::min(raw[i], dist[i]);
      for (int i = 0; i < N; ++i) s += tdist2[i];
      if (s < best_sum) {
        best_sum = s;
        best_center = ci;
        std::swap(tdist, tdist2);
      }
    }
    if (best_center < 0) best_center = 0;  // only with NaN input (OpenCV raises StsNoConv)
    chosen[k] = best_center;
    sum0 = best_sum;
    std::swap(dist, tdist);
  }
  for (int k = 0; k < K; ++k) std::memcpy(out_centers + (size_t)k * dims, data + (size_t)chosen[k] * dims, dims * sizeof(float));
}

double KMeans(const float* data, int N, int dims, int K, int max_iter, double eps, int attempts, CvRng& rng,
              int32_t* best_labels, float* centers_out) {
  attempts = std::max(attempts, 1);
  eps = std::max(eps, 0.0);
  eps *= eps;
  int max_count = std::min(std::max(max_iter, 2), 100);
  if (K == 1) {
    attempts = 1;
    max_count = 2;
  }
  std::vector<float> centers((size_t)K * dims), old_centers((size_t)K * dims), temp(dims), ppbuf;
  std::vector<int> counters(K), chosen;
  std::vector<int32_t> labels(N);
  std::vector<double> dists(N);
  Soa soa;
  soa.Load(data, N, dims);
  std::vector<float> dk((size_t)K * N);  // distances to centre k, all samples
  double best_compactness = DBL_MAX;
  for (int a = 0; a < attempts; ++a) {
    double compactness = 0;
    for (int iter = 0;;) {
      double max_center_shift = iter == 0 ? DBL_MAX : 0.0;
      centers.swap(old_centers);
      if (iter == 0) {
        CentersPP(data, soa, N, dims, K, rng, centers.data(), ppbuf, chosen);
      } else {
        std::fill(centers.begin(), centers.end(), 0.f);
        std::fill(counters.begin(), counters.end(), 0);
        for (int i = 0; i < N; ++i) {
          const float* s = data + (size_t)i * dims;
          float* c = centers.data() + (size_t)labels[i] * dims;
          for (int j = 0; j < dims; ++j) c[j] += s[j];
          counters[labels[i]]++;
        }
        for (int k = 0; k < K; ++k) {
          if (counters[k] != 0) continue;
          // empty cluster: split off the point of the biggest cluster farthest from its centre
          int max_k = 0;
          for (int k1 = 1; k1 < K; ++k1)
            if (counters[max_k] < counters[k1]) max_k = k1;
          double max_dist = 0;
          int farthest = -1;
          float* base = centers.data() + (size_t)max_k * dims;
          const float scale = 1.f / counters[max_k];
          for (int j = 0; j < dims; ++j) temp[j] = base[j] * scale;
          for (int i = 0; i < N; ++i) {
            if (labels[i] != max_k) continue;
            const double d = NormL2Sqr(data + (size_t)i * dims, temp.data(), dims);
            if (max_dist <= d) {
              max_dist = d;
              farthest = i;
            }
          }
          counters[max_k]--;
          counters[k]++;
          labels[farthest] = k;
          const float* s = data + (size_t)farthest * dims;
          float* cc = centers.data() + (size_t)k * dims;
          for (int j = 0; j < dims; ++j) {
            base[j] -= s[j];
            cc[j] += s[j];
          }
        }
        for (int k = 0; k < K; ++k) {
          float* c = centers.data() + (size_t)k * dims;
          const float scale = 1.f / counters[k];
          for (int j = 0; j < dims; ++j) c[j] *= scale;
          if (iter > 0) {
            double d = 0;
            const float* oc = old_centers.data() + (size_t)k * dims;
            for (int j = 0; j < dims; ++j) {
              const double t = c[j] - oc[j];
              d += t * t;
            }
            max_center_shift = std::max(max_center_shift, d);
          }
        }
      }
      const bool last = (++iter == std::max(max_count, 2) || max_center_shift <= eps);
      if (last) {
        // labels are not re-assigned on the last pass (no new empty clusters); distances only
        compactness = 0;
        for (int i = 0; i < N; ++i) {
          dists[i] = NormL2Sqr(data + (size_t)i * dims, centers.data() + (size_t)labels[i] * dims, dims);
          compactness += dists[i];
        }
        break;
      }
      // assignment: nearest centre, the first one among equals (strict ">" in centre order)
      for (int k = 0; k < K; ++k) DistToPoint(soa, centers.data() + (size_t)k * dims, dk.data() + (size_t)k * N);
      for (int i = 0; i < N; ++i) {
        int kb = 0;
        float md = dk[i];
        for (int k = 1; k < K; ++k) {
          const float d = dk[(size_t)k * N + i];
          if (md > d) {
            md = d;
            kb = k;
          }
        }
        dists[i] = md;
        labels[i] = kb;
      }
    }
    if (compactness < best_compactness) {
      best_compactness = compactness;
      if (centers_out) std::memcpy(centers_out, centers.data(), centers.size() * sizeof(float));
      std::memcpy(best_labels, labels.data(), (size_t)N * sizeof(int32_t));
    }
  }
  return best_compactness;
}

// ---- the chain (libs/encoder.cpp:491-624) ---------------------------------------------------------
const char* ValidateSegmentConfig(const SegmentConfig& c) {
  // libs/encoder.cpp:20-60
  if (c.ransac.inlier_thresh < 0) return "invalid inlier threshold: must be >= 0";
  if (c.ransac.success_prob < 0) return "invalid success probability: must be >= 0";
  if (c.ransac.inlier_ratio < 0) return "invalid inlier ratio: must be >= 0";
  if (c.kmeans.cluster_count == 0) return "invalid cluster count: must be > 0";
  if (c.kmeans.attempt_count == 0) return "invalid attempt count: must be > 0";
  if (c.kmeans.max_iter_count == 0) return "invalid maximum iteration count: must be > 0";
  if (c.kmeans.epsilon <= 0) return "invalid epsilon: must be > 0";
  // libs/encoder.cpp:96-101
  if (c.connected_components_connectivity != 4 && c.connected_components_connectivity != 8)
    return "invalid connected components connectivity: must be either 4 or 8";
  // not checked by the reference (cv::getStructuringElement would assert)
  if (c.morph_rect_w == 0 || c.morph_rect_h == 0) return "invalid morphology rectangle: must be > 0";
  if (c.ransac.subset_sz == 0) return "invalid subset size: must be > 0";
  return "";
}

MotionSegmenter::MotionSegmenter(const SegmentConfig& cfg, uint w, uint h) : cfg_(cfg), w_(w), h_(h) {
  const size_t n = (size_t)w * h;
  // the reference asserts motion_field_sz >= subset_sz (libs/motion.cpp:196); with fewer vectors its
  // distinct-subset draw never terminates
  if (n < cfg.ransac.subset_sz || cfg.ransac.subset_sz == 0)
    throw Error(1, "motion field smaller than the RANSAC subset size");
  mask_.resize(n);
  cluster_mask_.resize(n);
  comp_ids_.resize(n);
}

void MotionSegmenter::operator()(const Vec2f* mv, MinstdRand0& ransac_rng, CvRng& kmeans_rng, uint* types,
                                 Vec2f* global_motion) {
  const uint n = w_ * h_;
  Vec2f gm{0.f, 0.f};
  float rmse = 0;
  RansacGlobalMotion(mv, n, cfg_.ransac, ransac_rng, &rmse, &gm, &inliers_);  // libs/encoder.cpp:491-498
  if (global_motion) *global_motion = gm;
  std::fill(mask_.begin(), mask_.end(), (uint8_t)255);                         // :507-513
  for (uint i : inliers_) mask_[i] = 0;
  MorphologyEx(mask_.data(), w_, h_, kClose, cfg_.morph_rect_w, cfg_.morph_rect_h);  // :519-522
  MorphologyEx(mask_.data(), w_, h_, kOpen, cfg_.morph_rect_w, cfg_.morph_rect_h);
  fg_.clear();                                                                 // :534-545
  for (uint i = 0; i < n; ++i)
    if (mask_[i] == 255) fg_.push_back(i);
  std::fill(types, types + n, 0u);                                             // :547-549
  if (fg_.empty()) return;
  const uint K = std::min<uint>(cfg_.kmeans.cluster_count, (uint)fg_.size());  // :556-557
  // BuildMvFeatures (libs/encoder.cpp:296-321): through Vec4f::operator[] (libs/math.hpp:285-291,
  // indexing from &x) the four floats handed to cv::kmeans are (0, mv.x, block x, block y) -- mv.y
  // is overwritten by the block position.  Kept, since it decides the clusters.
  features_.resize((size_t)fg_.size() * 4);
  for (size_t i = 0; i < fg_.size(); ++i) {
    const uint j = fg_[i];
    features_[4 * i + 0] = 0.f;
    features_[4 * i + 1] = mv[j].x;
    features_[4 * i + 2] = (float)((j % w_) * cfg_.mv_block_w);
    features_[4 * i + 3] = (float)((j / w_) * cfg_.mv_block_h);
  }
  cluster_ids_.resize(fg_.size());
  KMeans(features_.data(), (int)fg_.size(), 4, (int)K, (int)cfg_.kmeans.max_iter_count, cfg_.kmeans.epsilon,
         (int)cfg_.kmeans.attempt_count, kmeans_rng, cluster_ids_.data(), nullptr);  // :566-575
  uint offset = 0;  // BLOCK_TYPE_BACKGROUND
  for (uint cid = 0; cid < K; ++cid) {                                         // :597-624
    std::fill(cluster_mask_.begin(), cluster_mask_.end(), (uint8_t)0);
    for (size_t i = 0; i < fg_.size(); ++i)
      if ((uint)cluster_ids_[i] == cid) cluster_mask_[fg_[i]] = 255;
    const uint n_labels = ConnectedComponents(cluster_mask_.data(), w_, h_, cfg_.connected_components_connectivity,
                                              comp_ids_.data());
    for (uint i : fg_)
      if (comp_ids_[i] != 0) types[i] = (uint)comp_ids_[i] + offset;
    offset += n_labels;  // the count includes the background label, as in the reference
  }
}

// ---- batch stage -----------------------------------------------------------------------------------
static uint64_t SplitMix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}

void BlockTypeStage::FrameGenerators(uint64_t seed, uint64_t frame, MinstdRand0* r, CvRng* k) {
  const uint64_t a = SplitMix64(seed ^ SplitMix64(frame));
  r->Seed((uint32_t)(a >> 32));
  *k = CvRng((SplitMix64(a) & 0x7fffffffull) | 1ull);  // a value cv::setRNGSeed(int) can express
}

BlockTypeStage::BlockTypeStage(const SegmentConfig& cfg, uint w, uint h, uint64_t seed, uint threads)
    : cfg_(cfg), w_(w), h_(h), threads_(threads), seed_(seed) {
  if (seed_ == 0) {
    std::random_device rdev;
    seed_ = ((uint64_t)rdev() << 32) | rdev();
    if (seed_ == 0) seed_ = 1;
  }
  if (threads_ == 0) threads_ = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  for (uint t = 0; t < threads_; ++t) workers_.emplace_back(cfg_, w_, h_);
}

void BlockTypeStage::Run(const Vec2f* mv, uint n, uint64_t first_frame, uint* types) {
  const size_t mvn = (size_t)w_ * h_;
  std::atomic<uint> next{0};
  auto work = [&](uint t) {
    for (uint i = next.fetch_add(1); i < n; i = next.fetch_add(1)) {
      MinstdRand0 r;
      CvRng k;
      FrameGenerators(seed_, first_frame + i, &r, &k);
      workers_[t](mv + i * mvn, r, k, types + i * mvn);
    }
  };
  const uint nt = std::min(threads_, n);
  if (nt <= 1) {
    work(0);
    return;
  }
  std::vector<std::thread> pool;
  for (uint t = 1; t < nt; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
}

}  // namespace svc

// ---- reference signatures ---------------------------------------------------------------------------
Vec2f EstimateGlobalMotionAvg(const Vec2f* mf, uint sz) {  // libs/motion.cpp:45-53 (running mean)
  Vec2f avg{0.f, 0.f};
  for (uint i = 0; i < sz; ++i) {
    const float s = 1.0f / (i + 1);
    avg.x += (mf[i].x - avg.x) * s;
    avg.y += (mf[i].y - avg.y) * s;
  }
  return avg;
}

namespace {
std::mutex g_rng_mutex;
bool g_rng_seeded = false;
svc::MinstdRand0 g_rng;
}  // namespace

void svc::SeedGlobalMotionRng(uint32_t seed) {
  std::lock_guard<std::mutex> l(g_rng_mutex);
  g_rng.Seed(seed);
  g_rng_seeded = true;
}

void EstimateGlobalMotionRansac(const Vec2f* mf, uint n, RansacParams params, float* rmse, Vec2f* gm,
                                std::vector<uint>* inliers) {
  std::lock_guard<std::mutex> l(g_rng_mutex);
  if (!g_rng_seeded) {  // libs/motion.cpp:186-187
    std::random_device rdev;
    g_rng.Seed(rdev());
    g_rng_seeded = true;
  }
  svc::RansacGlobalMotion(mf, n, params, g_rng, rmse, gm, inliers);
}

// ---- C ABI (include/svc_segment.h) ------------------------------------------------------------------
namespace {
thread_local std::string g_seg_err;
int seg_fail(const char* m) {
  g_seg_err = m;
  return 1;
}
svc::SegmentConfig to_cfg(const svc_seg_config& c) {
  svc::SegmentConfig s;
  s.ransac = RansacParams{c.ransac_subset_sz, c.ransac_inlier_thresh, c.ransac_success_prob, c.ransac_inlier_ratio};
  s.morph_rect_w = c.morph_rect_w;
  s.morph_rect_h = c.morph_rect_h;
  s.kmeans = KMeansParams{c.kmeans_cluster_count, c.kmeans_attempt_count, c.kmeans_max_iter_count, c.kmeans_epsilon};
  s.connected_components_connectivity = c.connected_components_connectivity;
  s.mv_block_w = c.mv_block_w;
  s.mv_block_h = c.mv_block_h;
  return s;
}
}  // namespace

#define SVC_SEG_API extern "C" __attribute__((visibility("default")))

SVC_SEG_API const char* svc_seg_last_error(void) { return g_seg_err.c_str(); }

SVC_SEG_API void svc_seg_default_config(svc_seg_config* c) {
  if (!c) return;
  const svc::SegmentConfig d;
  *c = svc_seg_config{d.ransac.subset_sz, d.ransac.inlier_thresh, d.ransac.success_prob, d.ransac.inlier_ratio,
                      d.morph_rect_w, d.morph_rect_h, d.kmeans.cluster_count, d.kmeans.attempt_count,
                      d.kmeans.max_iter_count, d.kmeans.epsilon, d.connected_components_connectivity,
                      d.mv_block_w, d.mv_block_h};
}

SVC_SEG_API int svc_seg_validate(const svc_seg_config* c) {
  if (!c) return seg_fail("null config");
  const char* m = svc::ValidateSegmentConfig(to_cfg(*c));
  return *m ? seg_fail(m) : 0;
}

SVC_SEG_API int svc_seg_ransac(const float* mv_xy, uint32_t n, uint32_t subset_sz, float inlier_thresh,
                               float success_prob, float inlier_ratio, uint32_t* rng_state, float* rmse,
                               float* gm_xy, uint32_t* inliers, uint32_t* n_inliers) {
  if (!mv_xy || !rng_state || !rmse || !gm_xy || !inliers || !n_inliers) return seg_fail("null argument");
  if (subset_sz == 0 || n < subset_sz) return seg_fail("motion field smaller than the subset size");
  svc::MinstdRand0 rng(*rng_state);
  std::vector<uint> in;
  Vec2f gm{gm_xy[0], gm_xy[1]};
  svc::RansacGlobalMotion(reinterpret_cast<const Vec2f*>(mv_xy), n,
                          RansacParams{subset_sz, inlier_thresh, success_prob, inlier_ratio}, rng, rmse, &gm, &in);
  gm_xy[0] = gm.x;
  gm_xy[1] = gm.y;
  *n_inliers = (uint32_t)in.size();
  std::copy(in.begin(), in.end(), inliers);
  *rng_state = rng.state;
  return 0;
}

SVC_SEG_API int svc_seg_global_motion_avg(const float* mv_xy, uint32_t n, float* gm_xy) {
  if (!mv_xy || !gm_xy) return seg_fail("null argument");
  const Vec2f g = EstimateGlobalMotionAvg(reinterpret_cast<const Vec2f*>(mv_xy), n);
  gm_xy[0] = g.x;
  gm_xy[1] = g.y;
  return 0;
}

SVC_SEG_API int svc_seg_morphology(uint8_t* mask, uint32_t w, uint32_t h, uint32_t op, uint32_t rw, uint32_t rh) {
  if (!mask || !w || !h) return seg_fail("empty mask");
  if (op > 3 || !rw || !rh) return seg_fail("invalid morphology operation or rectangle");
  svc::MorphologyEx(mask, w, h, op, rw, rh);
  return 0;
}

SVC_SEG_API int svc_seg_connected_components(const uint8_t* mask, uint32_t w, uint32_t h, uint32_t connectivity,
                                             int32_t* labels, uint32_t* n_labels) {
  if (!mask || !labels || !n_labels || !w || !h) return seg_fail("null argument");
  if (connectivity != 4 && connectivity != 8)
    return seg_fail("invalid connected components connectivity: must be either 4 or 8");
  *n_labels = svc::ConnectedComponents(mask, w, h, connectivity, labels);
  return 0;
}

SVC_SEG_API int svc_seg_kmeans(const float* data, uint32_t n, uint32_t dims, uint32_t k, uint32_t max_iter, float eps,
                               uint32_t attempts, uint64_t* rng_state, int32_t* labels, float* centers,
                               double* compactness) {
  if (!data || !rng_state || !labels) return seg_fail("null argument");
  if (!k || !dims || n < k) return seg_fail("number of clusters must be > 0 and <= number of samples");
  svc::CvRng rng(*rng_state);
  const double c = svc::KMeans(data, (int)n, (int)dims, (int)k, (int)max_iter, eps, (int)attempts, rng, labels, centers);
  if (compactness) *compactness = c;
  *rng_state = rng.state;
  return 0;
}

SVC_SEG_API int svc_seg_block_types(const float* mv_xy, uint32_t w, uint32_t h, const svc_seg_config* cfg,
                                    uint32_t* ransac_rng_state, uint64_t* kmeans_rng_state, uint32_t* block_types,
                                    float* gm_xy) {
  if (!mv_xy || !cfg || !ransac_rng_state || !kmeans_rng_state || !block_types || !w || !h)
    return seg_fail("null argument");
  const svc::SegmentConfig c = to_cfg(*cfg);
  const char* m = svc::ValidateSegmentConfig(c);
  if (*m) return seg_fail(m);
  if ((uint64_t)w * h < c.ransac.subset_sz) return seg_fail("motion field smaller than the subset size");
  svc::MotionSegmenter seg(c, w, h);
  svc::MinstdRand0 r(*ransac_rng_state);
  svc::CvRng kr(*kmeans_rng_state);
  Vec2f gm{0.f, 0.f};
  seg(reinterpret_cast<const Vec2f*>(mv_xy), r, kr, block_types, &gm);
  if (gm_xy) {
    gm_xy[0] = gm.x;
    gm_xy[1] = gm.y;
  }
  *ransac_rng_state = r.state;
  *kmeans_rng_state = kr.state;
  return 0;
}

SVC_SEG_API void svc_seg_frame_generators(uint64_t seed, uint64_t frame, uint32_t* ransac_rng_state,
                                          uint64_t* kmeans_rng_state) {
  svc::MinstdRand0 r;
  svc::CvRng k;
  svc::BlockTypeStage::FrameGenerators(seed, frame, &r, &k);
  if (ransac_rng_state) *ransac_rng_state = r.state;
  if (kmeans_rng_state) *kmeans_rng_state = k.state;
}

SVC_SEG_API int svc_seg_block_types_batch(const float* mv_xy, uint32_t n, uint32_t w, uint32_t h,
                                          const svc_seg_config* cfg, uint64_t seed, uint64_t first_frame,
                                          uint32_t threads, uint32_t* block_types) {
  if (!mv_xy || !cfg || !block_types || !w || !h) return seg_fail("null argument");
  if (seed == 0) return seg_fail("seed must be non-zero (0 means 'random' in svc::Encoder)");
  const svc::SegmentConfig c = to_cfg(*cfg);
  const char* m = svc::ValidateSegmentConfig(c);
  if (*m) return seg_fail(m);
  if ((uint64_t)w * h < c.ransac.subset_sz) return seg_fail("motion field smaller than the subset size");
  svc::BlockTypeStage stage(c, w, h, seed, threads);
  stage.Run(reinterpret_cast<const Vec2f*>(mv_xy), n, first_frame, block_types);
  return 0;
}
