/* rcc_ceres_adapter.h -- header-only Ceres front end of the B200 bundle-adjustment path.
 *
 * What it is for: the reference's missing optimiser stage (between real_preprocessing/src/camera_pose.cpp and
 * opt_visualization.cpp:46-66,118-138) would build a ceres::Problem with one reprojection residual block per
 * detected tag and call ceres::Solve.  With this header that code keeps its shape -- AddResidualBlock per tag,
 * SetParameterBlockConstant for the world tag (camera_pose.cpp:71-80), ceres::Solve -- while every residual and
 * Jacobian is computed in ONE batched GPU evaluation per Ceres evaluation point:
 *
 *   rcc_ceres::GpuBatch batch(intr4, dist5);                          // K / distortion, camera_pose.cpp:38-39,55-68
 *   for each detection:
 *     problem.AddResidualBlock(batch.AddTag(view6, marker6, size, pixels8), nullptr, intr4, dist5, view6, marker6);
 *   batch.Finalize(0);                                                // device ordinal: uploads the observations
 *   options.evaluation_callback = &batch;                             // Solver::Options (1.14) / Problem::Options (2.x)
 *
 * Contract relied upon (Ceres public API, 1.14 and later; not in /root/reference, which never calls Ceres):
 *   bool CostFunction::Evaluate(double const* const* parameters, double* residuals, double** jacobians) const
 *     jacobians may be NULL; jacobians[i] may be NULL; jacobians[i][r * block_size_i + c] = d res_r / d param_i,c
 *   void EvaluationCallback::PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point)
 *     called once per evaluation, after Ceres has written the evaluation point into the USER's parameter arrays
 *     and before any Evaluate; Evaluate may then run from several threads -> it only reads.
 * rcc_ba_evaluate already emits the per-block row-major layout, so Evaluate is five memcpys.
 *
 * Compiles against real Ceres (<ceres/ceres.h>) or against tests/ceres_mock (the same declarations, nothing else),
 * which is how tests/test_ceres_adapter.py exercises it in a container without Ceres.
 */
#ifndef RCC_CERES_ADAPTER_H
#define RCC_CERES_ADAPTER_H

#include <ceres/ceres.h>

#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "rcc_ba.h"

namespace rcc_ceres {

class GpuBatch;

/* one tag in one frame: 8 residuals (u0 v0 .. u3 v3, corner_detections.cpp:34-37) against
 * intr[4], dist[5], view[6] (world_T_camera), marker[6] (world_T_target) */
class TagReprojection : public ceres::SizedCostFunction<8, 4, 5, 6, 6> {
 public:
  TagReprojection(const GpuBatch* batch, int64_t index) : batch_(batch), n_(index) {}
  bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override;

 private:
  const GpuBatch* batch_;
  int64_t n_;
};

class GpuBatch : public ceres::EvaluationCallback {
 public:
  GpuBatch(double* intr4, double* dist5) : intr_(intr4), dist_(dist5) {}
  ~GpuBatch() override { rcc_ba_destroy(p_); }
  GpuBatch(const GpuBatch&) = delete;
  GpuBatch& operator=(const GpuBatch&) = delete;

  /* Register one detection; the returned cost function is owned by the ceres::Problem it is added to (Ceres's
   * default ownership).  view6 / marker6 are the caller's parameter blocks: blocks are identified by address. */
  ceres::CostFunction* AddTag(double* view6, double* marker6, double tag_size, const double pixels8[8]) {
    if (p_) throw std::logic_error("GpuBatch: AddTag after Finalize");
    view_idx_.push_back(Intern(view6, &view_of_, &views_));
    const int m = Intern(marker6, &marker_of_, &markers_);
    if ((size_t)m == sizes_.size()) sizes_.push_back(tag_size);
    else if (sizes_[m] != tag_size) throw std::invalid_argument("GpuBatch: one marker with two tag sizes");
    marker_idx_.push_back(m);
    pixels_.insert(pixels_.end(), pixels8, pixels8 + 8);
    return new TagReprojection(this, (int64_t)view_idx_.size() - 1);
  }

  /* Create the GPU problem.  Call once, after the last AddTag and before the first evaluation. */
  void Finalize(int device = 0) {
    if (p_) throw std::logic_error("GpuBatch: Finalize called twice");
    const int64_t n = (int64_t)view_idx_.size();
    rcc_ba_options opt = {RCC_MODEL_SINGLE, (int32_t)views_.size(), (int32_t)markers_.size(), 1, n, device, RCC_ELIM_AUTO};
    Check(rcc_ba_create(&opt, &p_), "rcc_ba_create");
    Check(rcc_ba_set_marker_sizes(p_, sizes_.data()), "rcc_ba_set_marker_sizes");
    Check(rcc_ba_set_observations(p_, view_idx_.data(), marker_idx_.data(), nullptr, pixels_.data()),
          "rcc_ba_set_observations");
    vbuf_.resize(views_.size() * 6);
    mbuf_.resize(markers_.size() * 6);
    res_.resize((size_t)n * 8);
    ji_.resize((size_t)n * 32);
    jd_.resize((size_t)n * 40);
    jv_.resize((size_t)n * 48);
    jm_.resize((size_t)n * 48);
  }

  /* ceres::EvaluationCallback: one batched GPU evaluation at the point Ceres has just written into the user's arrays */
  void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) override {
    if (!p_) throw std::logic_error("GpuBatch: PrepareForEvaluation before Finalize");
    if (!new_evaluation_point && (have_jac_ || !evaluate_jacobians)) return;   // same point, nothing new asked for
    for (size_t i = 0; i < views_.size(); ++i) std::memcpy(&vbuf_[6 * i], views_[i], 6 * sizeof(double));
    for (size_t i = 0; i < markers_.size(); ++i) std::memcpy(&mbuf_[6 * i], markers_[i], 6 * sizeof(double));
    Check(rcc_ba_set_intrinsics(p_, intr_, dist_), "rcc_ba_set_intrinsics");
    Check(rcc_ba_set_view_poses(p_, vbuf_.data()), "rcc_ba_set_view_poses");
    Check(rcc_ba_set_marker_poses(p_, mbuf_.data()), "rcc_ba_set_marker_poses");
    const int rc = rcc_ba_evaluate(p_, evaluate_jacobians ? 1 : 0, &cost_, res_.data(),
                                   evaluate_jacobians ? ji_.data() : nullptr, evaluate_jacobians ? jd_.data() : nullptr,
                                   evaluate_jacobians ? jv_.data() : nullptr, evaluate_jacobians ? jm_.data() : nullptr,
                                   nullptr);
    if (rc != RCC_OK && rc != RCC_EVAL_FAILED) Check(rc, "rcc_ba_evaluate");
    ok_ = rc == RCC_OK;            // RCC_EVAL_FAILED == a corner behind the camera: every Evaluate returns false
    have_jac_ = evaluate_jacobians;
    ++evaluations_;
  }

  int64_t num_tags() const { return (int64_t)view_idx_.size(); }
  int64_t evaluations() const { return evaluations_; }
  double cost() const { return cost_; }           /* 0.5 * sum r^2 of the last evaluation */
  rcc_ba_problem* handle() const { return p_; }

 private:
  friend class TagReprojection;
  static int Intern(double* ptr, std::map<double*, int>* index, std::vector<double*>* list) {
    auto it = index->find(ptr);
    if (it != index->end()) return it->second;
    const int i = (int)list->size();
    (*index)[ptr] = i;
    list->push_back(ptr);
    return i;
  }
  void Check(int rc, const char* what) const {
    if (rc != RCC_OK) throw std::runtime_error(std::string(what) + ": " + rcc_ba_last_error(p_));
  }

  double* intr_;
  double* dist_;
  std::map<double*, int> view_of_, marker_of_;
  std::vector<double*> views_, markers_;
  std::vector<int32_t> view_idx_, marker_idx_;
  std::vector<double> sizes_, pixels_, vbuf_, mbuf_;
  std::vector<double> res_, ji_, jd_, jv_, jm_;
  rcc_ba_problem* p_ = nullptr;
  double cost_ = 0.0;
  bool ok_ = false, have_jac_ = false;
  int64_t evaluations_ = 0;
};

inline bool TagReprojection::Evaluate(double const* const* /*parameters*/, double* residuals, double** jacobians) const {
  const GpuBatch& b = *batch_;
  if (!b.ok_) return false;
  std::memcpy(residuals, b.res_.data() + 8 * n_, 8 * sizeof(double));
  if (jacobians) {
    if (!b.have_jac_) return false;     // Ceres asked PrepareForEvaluation for residuals only
    if (jacobians[0]) std::memcpy(jacobians[0], b.ji_.data() + 32 * n_, 32 * sizeof(double));
    if (jacobians[1]) std::memcpy(jacobians[1], b.jd_.data() + 40 * n_, 40 * sizeof(double));
    if (jacobians[2]) std::memcpy(jacobians[2], b.jv_.data() + 48 * n_, 48 * sizeof(double));
    if (jacobians[3]) std::memcpy(jacobians[3], b.jm_.data() + 48 * n_, 48 * sizeof(double));
  }
  return true;
}

}  // namespace rcc_ceres

#endif /* RCC_CERES_ADAPTER_H */
