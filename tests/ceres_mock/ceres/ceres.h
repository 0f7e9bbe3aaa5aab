// TEST-ONLY stand-in for <ceres/ceres.h>: the declarations of the published Ceres Solver interface that
// include/rcc_ceres_adapter.h touches, with exactly the published signatures, and nothing of Ceres's
// implementation (Ceres is not installed in this container and is not part of /root/reference).
//   CostFunction / SizedCostFunction   ceres/cost_function.h, ceres/sized_cost_function.h
//   EvaluationCallback                 ceres/evaluation_callback.h (>= 1.14)
//   Problem::AddResidualBlock / SetParameterBlockConstant / Evaluate-like walk over the blocks
// The mock Problem only records what it is given and can walk its residual blocks the way Ceres's evaluator does
// (PrepareForEvaluation once, then Evaluate per block with per-block Jacobian pointers, NULL for constant blocks).
#ifndef CERES_MOCK_CERES_H
#define CERES_MOCK_CERES_H

#include <cstdint>
#include <set>
#include <vector>

namespace ceres {

class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  CostFunction(const CostFunction&) = delete;
  void operator=(const CostFunction&) = delete;
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int32_t>& parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int32_t>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int num_residuals) { num_residuals_ = num_residuals; }

 private:
  std::vector<int32_t> parameter_block_sizes_;
  int num_residuals_;
};

template <int kNumResiduals, int... Ns>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() {
    set_num_residuals(kNumResiduals);
    *mutable_parameter_block_sizes() = std::vector<int32_t>{Ns...};
  }
  virtual ~SizedCostFunction() {}
};

class EvaluationCallback {
 public:
  virtual ~EvaluationCallback() {}
  virtual void PrepareForEvaluation(bool evaluate_jacobians, bool new_evaluation_point) = 0;
};

class LossFunction;

class Problem {
 public:
  struct Block {
    CostFunction* cost;
    std::vector<double*> params;
  };
  ~Problem() {
    for (auto& b : blocks_) delete b.cost;      // Ceres's default: the problem owns its cost functions
  }
  template <typename... Ts>
  void* AddResidualBlock(CostFunction* cost_function, LossFunction* /*loss_function*/, double* x0, Ts*... xs) {
    blocks_.push_back(Block{cost_function, std::vector<double*>{x0, xs...}});
    return &blocks_.back();
  }
  void SetParameterBlockConstant(double* values) { constant_.insert(values); }
  int NumResidualBlocks() const { return (int)blocks_.size(); }

  // Walk the residual blocks like Ceres's evaluator.  residuals: sum of num_residuals; jacobians: per block, per
  // parameter block, row-major num_residuals x block_size, concatenated (zeros where the block is constant).
  bool Evaluate(EvaluationCallback* callback, bool evaluate_jacobians, bool new_point, std::vector<double>* residuals,
                std::vector<double>* jacobians) const {
    if (callback) callback->PrepareForEvaluation(evaluate_jacobians, new_point);
    residuals->clear();
    if (jacobians) jacobians->clear();
    for (const Block& b : blocks_) {
      const int nr = b.cost->num_residuals();
      const std::vector<int32_t>& sizes = b.cost->parameter_block_sizes();
      std::vector<double> r(nr);
      std::vector<std::vector<double>> J(sizes.size());
      std::vector<double*> jp(sizes.size(), nullptr);
      for (size_t i = 0; i < sizes.size(); ++i) {
        J[i].assign((size_t)nr * sizes[i], 0.0);
        if (!constant_.count(b.params[i])) jp[i] = J[i].data();
      }
      const bool ok = b.cost->Evaluate(b.params.data(), r.data(), evaluate_jacobians ? jp.data() : nullptr);
      if (!ok) return false;
      residuals->insert(residuals->end(), r.begin(), r.end());
      if (jacobians && evaluate_jacobians)
        for (auto& j : J) jacobians->insert(jacobians->end(), j.begin(), j.end());
    }
    return true;
  }

 private:
  std::vector<Block> blocks_;
  std::set<double*> constant_;
};

}  // namespace ceres

#endif  // CERES_MOCK_CERES_H
