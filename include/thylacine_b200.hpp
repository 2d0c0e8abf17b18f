// thylacine_b200.hpp -- header-only C++17 mirror of the reference's operator surface over the C ABI.
//
// The reference is a compiled (JVM) library whose toolchain is absent from this image; this is the host side a
// compiled caller links against: the reference's names, argument meaning and error behaviour, RAII around the opaque
// handles of thylacine_b200.h.  Paths cited are relative to /root/reference/src/main/scala/thylacine/.
//   model/components/prior/{GaussianPrior,UniformPrior,CauchyPrior}.scala
//   model/components/forwardmodel/{LinearForwardModel,NonLinearForwardModel}.scala
//   model/components/likelihood/{GaussianLinearLikelihood,GaussianLikelihood,CauchyLikelihood,UniformLikelihood}.scala
//   model/components/posterior/{UnnormalisedPosterior,GaussianAnalyticPosterior,HmcmcSampledPosterior,
//                               LeapfrogMcmcSampledPosterior,SlqIntegratedPosterior}.scala
//   config/{HmcmcConfig,LeapfrogMcmcConfig,SlqConfig}.scala
// What the reference asserts (AssertionError) or throws (RuntimeException) arrives as thylacine::Error carrying the
// thy_status.  A parameter collection is std::map<label, vector> (IndexedVectorCollection); labels are kept sorted,
// as the reference's flattening does (Posterior.scala:40-48).  No CPU fallback: without a GPU every evaluation throws
// Error{THY_ERR_NO_DEVICE}.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "thylacine_b200.h"

namespace thylacine {

using Vector = std::vector<double>;
using ModelParameters = std::map<std::string, Vector>;   // IndexedVectorCollection

struct Error : std::runtime_error {
  thy_status status;
  Error(thy_status s, const std::string& msg) : std::runtime_error(msg), status(s) {}
};
inline void check(thy_status s) {
  if (s != THY_OK) throw Error(s, thy_last_error());
}

// Row-major matrix (rows = observations), IndexedMatrixCollection block.
struct Matrix {
  int rows = 0, cols = 0;
  Vector data;
  Matrix() = default;
  Matrix(int r, int c, Vector d) : rows(r), cols(c), data(std::move(d)) {
    if ((size_t)r * c != data.size()) throw Error(THY_ERR_INVALID_ARGUMENT, "matrix size mismatch");
  }
  Matrix(const std::vector<Vector>& rowsv) {   // Vector[Vector[Double]]: outer index = row (MatrixContainer.scala:121-139)
    rows = (int)rowsv.size();
    cols = rows ? (int)rowsv[0].size() : 0;
    for (auto& r : rowsv) {
      if ((int)r.size() != cols) throw Error(THY_ERR_INVALID_ARGUMENT, "ragged matrix");
      data.insert(data.end(), r.begin(), r.end());
    }
  }
  static Matrix identity(int n) {
    Matrix m(n, n, Vector((size_t)n * n, 0.0));
    for (int i = 0; i < n; ++i) m.data[(size_t)i * n + i] = 1.0;
    return m;
  }
};

// ---------------------------------------------------------------------------------------------------- priors
struct Prior {
  enum Kind { GaussianCi, GaussianCov, Uniform, Cauchy } kind;
  std::string label;
  Vector a, b;   // values | maxBounds ; confidenceIntervals | covariance | minBounds
  int dimension() const { return (int)a.size(); }
};
struct GaussianPrior {
  // GaussianPrior.fromConfidenceIntervals (prior/GaussianPrior.scala:62-75)
  static Prior fromConfidenceIntervals(std::string label, Vector values, Vector confidenceIntervals) {
    if (values.size() != confidenceIntervals.size()) throw Error(THY_ERR_INVALID_ARGUMENT, "dimension mismatch");
    return {Prior::GaussianCi, std::move(label), std::move(values), std::move(confidenceIntervals)};
  }
  // GaussianPrior.fromCovarianceMatrix (:78-93)
  static Prior fromCovarianceMatrix(std::string label, Vector values, const Matrix& covarianceMatrix) {
    if (covarianceMatrix.rows != (int)values.size() || covarianceMatrix.cols != (int)values.size())
      throw Error(THY_ERR_INVALID_ARGUMENT, "dimension mismatch");
    return {Prior::GaussianCov, std::move(label), std::move(values), covarianceMatrix.data};
  }
};
struct UniformPrior {
  // UniformPrior.fromBounds(label, maxBounds, minBounds) -- max first (prior/UniformPrior.scala:68-77)
  static Prior fromBounds(std::string label, Vector maxBounds, Vector minBounds) {
    if (maxBounds.size() != minBounds.size()) throw Error(THY_ERR_INVALID_ARGUMENT, "dimension mismatch");
    return {Prior::Uniform, std::move(label), std::move(maxBounds), std::move(minBounds)};
  }
};
struct CauchyPrior {
  // CauchyPrior(label, values, confidenceIntervals) (prior/CauchyPrior.scala:50-63)
  static Prior apply(std::string label, Vector values, Vector confidenceIntervals) {
    if (values.size() != confidenceIntervals.size()) throw Error(THY_ERR_INVALID_ARGUMENT, "dimension mismatch");
    return {Prior::Cauchy, std::move(label), std::move(values), std::move(confidenceIntervals)};
  }
};

// ---------------------------------------------------------------------------------------------------- forward models
struct ForwardModel {
  std::map<std::string, Matrix> blocks;   // label -> coefficient block (sorted: LinearForwardModel.scala:87-101)
  bool hasOffset = false;
  Vector offset;
  int link = THY_LINK_IDENTITY, jacobian = THY_JAC_CONSTANT;
  double differential = 0.0;
  int rangeDimension() const { return blocks.empty() ? 0 : blocks.begin()->second.rows; }
};
struct LinearForwardModel {
  // LinearForwardModel.of(label, values, evalCacheDepth) (forwardmodel/LinearForwardModel.scala:117-131);
  // the cache depth is accepted and ignored (a per-point memo is meaningless for a batched device path)
  static ForwardModel of(const std::string& label, const Matrix& values, int /*evalCacheDepth*/ = 0) {
    ForwardModel f;
    f.blocks[label] = values;
    return f;
  }
  // LinearForwardModel.of(transform, vectorOffset, ...) (:76-115)
  static ForwardModel of(const std::map<std::string, Matrix>& transform, const Vector* vectorOffset = nullptr) {
    ForwardModel f;
    f.blocks = transform;
    int rows = -1;
    for (auto& kv : f.blocks) {
      if (rows >= 0 && kv.second.rows != rows)
        throw Error(THY_ERR_INVALID_ARGUMENT, "all blocks must share the range dimension (LinearForwardModel.scala:47)");
      rows = kv.second.rows;
    }
    if (vectorOffset) {
      f.hasOffset = true;
      f.offset = *vectorOffset;
    }
    return f;
  }
  // LinearForwardModel.identityOf(label, dimension, ...) (:150-160)
  static ForwardModel identityOf(const std::string& label, int dimension) { return of(label, Matrix::identity(dimension)); }
};
struct NonLinearForwardModel {
  // Device-registered family f(theta) = link(A . theta + offset) in place of the reference's JVM closure
  // (forwardmodel/NonLinearForwardModel.scala:44-85).  differential > 0 => the finite-difference Jacobian of
  // core/computation/FiniteDifferenceJacobian.scala:30-55; differential == 0 => analytic Jacobian (second overload).
  static ForwardModel of(thy_link link, const std::map<std::string, Matrix>& coefficients, double differential = 0.0,
                         const Vector* vectorOffset = nullptr) {
    ForwardModel f = LinearForwardModel::of(coefficients, vectorOffset);
    f.link = link;
    f.jacobian = differential > 0 ? THY_JAC_FINITE_DIFFERENCE : THY_JAC_ANALYTIC;
    f.differential = differential;
    return f;
  }
};

// ---------------------------------------------------------------------------------------------------- likelihoods
struct Likelihood {
  ForwardModel forwardModel;
  int distribution = THY_DIST_GAUSSIAN;
  Vector measurements, uncertainties;   // Uniform: upperBounds, lowerBounds
};
struct GaussianLikelihood {
  // GaussianLikelihood(forwardModel, measurements, uncertainties) (likelihood/GaussianLikelihood.scala:55-67)
  static Likelihood apply(ForwardModel fm, Vector measurements, Vector uncertainties) {
    if (fm.rangeDimension() != (int)measurements.size() || measurements.size() != uncertainties.size())
      throw Error(THY_ERR_INVALID_ARGUMENT, "range dimension mismatch (GaussianLinearLikelihood.scala:43-45)");
    return {std::move(fm), THY_DIST_GAUSSIAN, std::move(measurements), std::move(uncertainties)};
  }
};
struct GaussianLinearLikelihood {
  // GaussianLinearLikelihood.of(coefficients, measurements, uncertainties, priorLabel, evalCacheDepth) (:62-83)
  static Likelihood of(const Matrix& coefficients, Vector measurements, Vector uncertainties, const std::string& priorLabel,
                       int /*evalCacheDepth*/ = 0) {
    return GaussianLikelihood::apply(LinearForwardModel::of(priorLabel, coefficients), std::move(measurements),
                                     std::move(uncertainties));
  }
};
struct CauchyLikelihood {
  // CauchyLikelihood(forwardModel, measurements, uncertainties) / .of (likelihood/CauchyLikelihood.scala:59-92)
  static Likelihood apply(ForwardModel fm, Vector measurements, Vector uncertainties) {
    Likelihood l = GaussianLikelihood::apply(std::move(fm), std::move(measurements), std::move(uncertainties));
    l.distribution = THY_DIST_CAUCHY;
    return l;
  }
  static Likelihood of(const Matrix& coefficients, Vector measurements, Vector uncertainties, const std::string& priorLabel) {
    return apply(LinearForwardModel::of(priorLabel, coefficients), std::move(measurements), std::move(uncertainties));
  }
};
struct UniformLikelihood {
  // UniformLikelihood(forwardModel, upperBounds, lowerBounds) (likelihood/UniformLikelihood.scala:63-73)
  static Likelihood apply(ForwardModel fm, Vector upperBounds, Vector lowerBounds) {
    if (fm.rangeDimension() != (int)upperBounds.size() || upperBounds.size() != lowerBounds.size())
      throw Error(THY_ERR_INVALID_ARGUMENT, "range dimension mismatch");
    return {std::move(fm), THY_DIST_UNIFORM, std::move(upperBounds), std::move(lowerBounds)};
  }
};

// ---------------------------------------------------------------------------------------------------- posterior
struct PosteriorOptions {
  bool cauchyReferenceSign = true;   // reproduce CauchyDistribution.scala:70-79; samplers use the correct sign
  bool fdIncremental = false;
  int kernelPath = 0;
};

class UnnormalisedPosterior {   // posterior/UnnormalisedPosterior.scala:28-47 + Posterior.scala:37-73
 public:
  UnnormalisedPosterior(const std::vector<Prior>& priors, const std::vector<Likelihood>& likelihoods = {},
                        const PosteriorOptions& opt = {}) {
    check(thy_model_create(&h_));
    try {
      for (auto& p : priors) {
        const int d = p.dimension();
        switch (p.kind) {
          case Prior::GaussianCi: check(thy_model_add_gaussian_prior_ci(h_, p.label.c_str(), d, p.a.data(), p.b.data())); break;
          case Prior::GaussianCov: check(thy_model_add_gaussian_prior_cov(h_, p.label.c_str(), d, p.a.data(), p.b.data())); break;
          case Prior::Uniform: check(thy_model_add_uniform_prior(h_, p.label.c_str(), d, p.a.data(), p.b.data())); break;
          case Prior::Cauchy: check(thy_model_add_cauchy_prior(h_, p.label.c_str(), d, p.a.data(), p.b.data())); break;
        }
      }
      for (auto& l : likelihoods) addLikelihood(l);
      check(thy_model_set_option(h_, THY_OPT_CAUCHY_REFERENCE_SIGN, opt.cauchyReferenceSign ? 1 : 0));
      check(thy_model_set_option(h_, THY_OPT_FD_INCREMENTAL, opt.fdIncremental ? 1 : 0));
      check(thy_model_set_option(h_, THY_OPT_KERNEL_PATH, opt.kernelPath));
      check(thy_model_finalize(h_));
      check(thy_model_dimension(h_, &d_));
      for (auto& p : priors) {
        int off = 0, w = 0;
        check(thy_model_label_layout(h_, p.label.c_str(), &off, &w));
        layout_[p.label] = {off, w};
      }
    } catch (...) {
      thy_model_destroy(h_);
      h_ = nullptr;
      throw;
    }
  }
  UnnormalisedPosterior(const UnnormalisedPosterior&) = delete;
  UnnormalisedPosterior& operator=(const UnnormalisedPosterior&) = delete;
  UnnormalisedPosterior(UnnormalisedPosterior&& o) noexcept : h_(o.h_), d_(o.d_), layout_(std::move(o.layout_)) { o.h_ = nullptr; }
  ~UnnormalisedPosterior() {
    if (h_) thy_model_destroy(h_);
  }

  int domainDimension() const { return d_; }                                  // Posterior.scala:37-38
  const std::map<std::string, std::pair<int, int>>& layout() const { return layout_; }
  thy_model_t handle() const { return h_; }

  // ModelParameterContext flatten / unflatten (core/values/modelparameters/ModelParameterContext.scala:42-80)
  Vector flatten(const ModelParameters& p) const {
    Vector x(d_, 0.0);
    for (auto& kv : p) {
      auto it = layout_.find(kv.first);
      if (it == layout_.end()) throw Error(THY_ERR_UNKNOWN_LABEL, "unknown label '" + kv.first + "'");
      if ((int)kv.second.size() != it->second.second) throw Error(THY_ERR_INVALID_ARGUMENT, "dimension mismatch for '" + kv.first + "'");
      for (int j = 0; j < it->second.second; ++j) x[it->second.first + j] = kv.second[j];
    }
    return x;
  }
  ModelParameters unflatten(const double* x) const {
    ModelParameters out;
    for (auto& kv : layout_) out[kv.first] = Vector(x + kv.second.first, x + kv.second.first + kv.second.second);
    return out;
  }

  // ModelParameterPdf.logPdfAt / logPdfGradientAt (ModelParameterPdf.scala:33-46,84-95)
  double logPdfAt(const ModelParameters& p) const {
    const Vector x = flatten(p);
    double lp = 0.0;
    check(thy_logpdf_batch(h_, 1, x.data(), &lp));
    return lp;
  }
  ModelParameters logPdfGradientAt(const ModelParameters& p) const {
    const Vector x = flatten(p);
    Vector g(d_);
    double lp = 0.0;
    check(thy_logpdf_grad_batch(h_, 1, x.data(), &lp, g.data()));
    return unflatten(g.data());
  }
  // batched surface: X is [B * d] chain-major
  Vector logPdfBatch(const Vector& X) const {
    const int64_t B = d_ ? (int64_t)X.size() / d_ : 0;
    Vector lp((size_t)B);
    check(thy_logpdf_batch(h_, B, X.data(), lp.data()));
    return lp;
  }
  std::pair<Vector, Vector> logPdfGradBatch(const Vector& X) const {
    const int64_t B = d_ ? (int64_t)X.size() / d_ : 0;
    Vector lp((size_t)B), g(X.size());
    check(thy_logpdf_grad_batch(h_, B, X.data(), lp.data(), g.data()));
    return {std::move(lp), std::move(g)};
  }
  // Posterior.samplePriors (Posterior.scala:64-65)
  Vector samplePriors(int64_t B, uint64_t rngSeed) const {
    Vector x((size_t)B * d_);
    check(thy_sample_priors(h_, B, rngSeed, x.data()));
    return x;
  }
  std::string lastKernelPath() const { return thy_model_last_kernel_path(h_); }

  // The same posterior with its GaussianLinearLikelihood terms folded onto d observations (thy_model_collapse).
  UnnormalisedPosterior collapsed() const {
    thy_model_t c = nullptr;
    check(thy_model_collapse(h_, &c));
    return UnnormalisedPosterior(c, d_, layout_);
  }

 private:
  UnnormalisedPosterior(thy_model_t h, int d, std::map<std::string, std::pair<int, int>> layout)
      : h_(h), d_(d), layout_(std::move(layout)) {}
  void addLikelihood(const Likelihood& l) {
    const ForwardModel& f = l.forwardModel;
    std::vector<const char*> labels;
    std::vector<int> dims;
    std::vector<const double*> blocks;
    for (auto& kv : f.blocks) {
      labels.push_back(kv.first.c_str());
      dims.push_back(kv.second.cols);
      blocks.push_back(kv.second.data.data());
    }
    thy_likelihood_desc d{};
    d.distribution = l.distribution;
    d.link = f.link;
    d.jacobian = f.jacobian;
    d.n = f.rangeDimension();
    d.nblocks = (int)labels.size();
    d.labels = labels.data();
    d.dims = dims.data();
    d.blocks = blocks.data();
    d.offset = f.hasOffset ? f.offset.data() : nullptr;
    d.measurements = l.measurements.data();
    d.uncertainties = l.uncertainties.data();
    d.differential = f.differential;
    check(thy_model_add_likelihood(h_, &d));
  }
  thy_model_t h_ = nullptr;
  int d_ = 0;
  std::map<std::string, std::pair<int, int>> layout_;
};

// GaussianAnalyticPosterior (posterior/GaussianAnalyticPosterior.scala:36-169): mean, covarianceStridedVector, maxLogPdf
class GaussianAnalyticPosterior {
 public:
  GaussianAnalyticPosterior(const std::vector<Prior>& priors, const std::vector<Likelihood>& likelihoods)
      : posterior_(priors, likelihoods) {
    for (auto& p : priors)
      if (p.kind != Prior::GaussianCi && p.kind != Prior::GaussianCov)
        throw Error(THY_ERR_INVALID_ARGUMENT, "GaussianAnalyticPosterior takes GaussianPriors (:39)");
    const int d = posterior_.domainDimension();
    mean_.resize(d);
    cov_.resize((size_t)d * d);
    check(thy_model_gaussian_posterior(posterior_.handle(), mean_.data(), cov_.data(), nullptr, &logEvidence_, &maxLogPdf_));
  }
  ModelParameters mean() const { return posterior_.unflatten(mean_.data()); }
  const Vector& covarianceStridedVector() const { return cov_; }
  double maxLogPdf() const { return maxLogPdf_; }
  double logEvidence() const { return logEvidence_; }   // not in the reference: the SLQ yardstick
  const UnnormalisedPosterior& posterior() const { return posterior_; }

 private:
  UnnormalisedPosterior posterior_;
  Vector mean_, cov_;
  double logEvidence_ = 0.0, maxLogPdf_ = 0.0;
};

// ---------------------------------------------------------------------------------------------------- HMC
struct HmcmcConfig {   // config/HmcmcConfig.scala:20-25
  int stepsBetweenSamples, stepsInDynamicsSimulation, warmupStepCount;
  double dynamicsSimulationStepSize;
};
struct HmcmcTelemetryUpdate {   // core/telemetry/HmcmcTelemetryUpdate.scala:20-25
  int64_t jumpAttempts, jumpAcceptances;
  double hamiltonianDifferential;
};
class HmcmcSampledPosterior {   // posterior/HmcmcSampledPosterior.scala:63-75 + sampling/hmcmc/HmcmcEngine.scala:55-214
 public:
  HmcmcSampledPosterior(const HmcmcConfig& cfg, const UnnormalisedPosterior& posterior, int64_t numberOfChains = 1,
                        uint64_t rngSeed = 0, const ModelParameters* seed = nullptr)
      : posterior_(posterior), chains_(numberOfChains) {
    thy_hmcmc_config c{cfg.stepsBetweenSamples, cfg.stepsInDynamicsSimulation, cfg.warmupStepCount, cfg.dynamicsSimulationStepSize};
    Vector seeds;
    if (seed) {
      const Vector x = posterior.flatten(*seed);
      for (int64_t b = 0; b < numberOfChains; ++b) seeds.insert(seeds.end(), x.begin(), x.end());
    }
    check(thy_hmc_create(posterior.handle(), &c, numberOfChains, rngSeed, seed ? seeds.data() : nullptr, &h_));
  }
  HmcmcSampledPosterior(const HmcmcSampledPosterior&) = delete;
  HmcmcSampledPosterior& operator=(const HmcmcSampledPosterior&) = delete;
  ~HmcmcSampledPosterior() {
    if (h_) thy_hmc_destroy(h_);
  }
  // ModelParameterSampler.sample(n) (sampling/ModelParameterSampler.scala:36-37): chain 0, as labelled collections
  std::vector<ModelParameters> sample(int numberOfSamples) {
    const Vector all = sampleChains(numberOfSamples);
    std::vector<ModelParameters> out;
    const int d = posterior_.domainDimension();
    for (int k = 0; k < numberOfSamples; ++k) out.push_back(posterior_.unflatten(all.data() + (size_t)k * d));
    return out;
  }
  // every chain: [chain][k][d]
  Vector sampleChains(int numberOfSamples, bool rerunBurnIn = true) {
    Vector out((size_t)chains_ * numberOfSamples * posterior_.domainDimension());
    check(thy_hmc_sample(h_, numberOfSamples, out.data(), rerunBurnIn ? 1 : 0));
    return out;
  }
  std::vector<HmcmcTelemetryUpdate> telemetry() const {
    std::vector<int64_t> att((size_t)chains_), acc((size_t)chains_);
    Vector dh((size_t)chains_);
    check(thy_hmc_stats(h_, att.data(), acc.data(), dh.data()));
    std::vector<HmcmcTelemetryUpdate> out;
    for (int64_t b = 0; b < chains_; ++b) out.push_back({att[b], acc[b], dh[b]});
    return out;
  }

 private:
  const UnnormalisedPosterior& posterior_;
  int64_t chains_;
  thy_hmc_t h_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------- Leapfrog MCMC
struct LeapfrogMcmcConfig {   // config/LeapfrogMcmcConfig.scala:20-24
  int stepsBetweenSamples, warmupStepCount, samplePoolSize;
};
class LeapfrogMcmcSampledPosterior {   // posterior/LeapfrogMcmcSampledPosterior.scala:72-106
 public:
  // `.of(config, distanceCalculation, posterior, ..., seed)`: the distance closure becomes a named thy_distance
  static LeapfrogMcmcSampledPosterior of(const LeapfrogMcmcConfig& cfg, thy_distance distanceCalculation,
                                         const UnnormalisedPosterior& posterior, uint64_t rngSeed = 0,
                                         const ModelParameters* seed = nullptr) {
    thy_leapfrog_mcmc_config c{cfg.stepsBetweenSamples, cfg.warmupStepCount, cfg.samplePoolSize};
    Vector x;
    if (seed) x = posterior.flatten(*seed);
    thy_lfm_t h = nullptr;
    check(thy_lfm_create(posterior.handle(), &c, distanceCalculation, rngSeed, seed ? x.data() : nullptr, &h));
    return LeapfrogMcmcSampledPosterior(posterior, h);
  }
  LeapfrogMcmcSampledPosterior(LeapfrogMcmcSampledPosterior&& o) noexcept : posterior_(o.posterior_), h_(o.h_) { o.h_ = nullptr; }
  LeapfrogMcmcSampledPosterior(const LeapfrogMcmcSampledPosterior&) = delete;
  ~LeapfrogMcmcSampledPosterior() {
    if (h_) thy_lfm_destroy(h_);
  }
  std::vector<ModelParameters> sample(int numberOfSamples) {
    const int d = posterior_.domainDimension();
    Vector out((size_t)numberOfSamples * d);
    check(thy_lfm_sample(h_, numberOfSamples, out.data()));
    std::vector<ModelParameters> res;
    for (int k = 0; k < numberOfSamples; ++k) res.push_back(posterior_.unflatten(out.data() + (size_t)k * d));
    return res;
  }

 private:
  LeapfrogMcmcSampledPosterior(const UnnormalisedPosterior& p, thy_lfm_t h) : posterior_(p), h_(h) {}
  const UnnormalisedPosterior& posterior_;
  thy_lfm_t h_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------- SLQ
struct SlqConfig {   // config/SlqConfig.scala:20-28
  int poolSize, abscissaNumber;
  double domainScalingIncrement, targetAcceptanceProbability;
  int sampleParallelism, maxIterationCount, minIterationCount;
};
class SlqIntegratedPosterior {   // posterior/SlqIntegratedPosterior.scala:76-125 + integration/slq/SlqEngine.scala:435-469
 public:
  static SlqIntegratedPosterior of(const SlqConfig& cfg, const UnnormalisedPosterior& posterior, uint64_t rngSeed = 0,
                                   int constrainedWalkSteps = 0, bool keepRetiredPoints = false, thy_comm_t comm = nullptr) {
    thy_slq_config c{};
    c.poolSize = cfg.poolSize;
    c.abscissaNumber = cfg.abscissaNumber;
    c.domainScalingIncrement = cfg.domainScalingIncrement;
    c.targetAcceptanceProbability = cfg.targetAcceptanceProbability;
    c.sampleParallelism = cfg.sampleParallelism;
    c.maxIterationCount = cfg.maxIterationCount;
    c.minIterationCount = cfg.minIterationCount;
    c.constrainedWalkSteps = constrainedWalkSteps;
    c.keepRetiredPoints = keepRetiredPoints ? 1 : 0;
    thy_slq_t h = nullptr;
    check(thy_slq_create(posterior.handle(), &c, rngSeed, nullptr, 0, comm, &h));
    return SlqIntegratedPosterior(posterior, h);
  }
  SlqIntegratedPosterior(SlqIntegratedPosterior&& o) noexcept : posterior_(o.posterior_), h_(o.h_), stats_(o.stats_) { o.h_ = nullptr; }
  SlqIntegratedPosterior(const SlqIntegratedPosterior&) = delete;
  ~SlqIntegratedPosterior() {
    if (h_) thy_slq_destroy(h_);
  }
  void rebuildSampleSimulation() { check(thy_slq_run(h_, &stats_)); }        // SlqIntegratedPosterior.scala:76-80
  double logEvidence() const { return stats_.log_evidence_mean; }
  const thy_slq_stats& stats() const { return stats_; }
  std::vector<ModelParameters> sample(int numberOfSamples, uint64_t rngSeed = 1) {   // SlqEngine.scala:468-469
    const int d = posterior_.domainDimension();
    Vector out((size_t)numberOfSamples * d);
    check(thy_slq_sample(h_, numberOfSamples, rngSeed, out.data()));
    std::vector<ModelParameters> res;
    for (int k = 0; k < numberOfSamples; ++k) res.push_back(posterior_.unflatten(out.data() + (size_t)k * d));
    return res;
  }

 private:
  SlqIntegratedPosterior(const UnnormalisedPosterior& p, thy_slq_t h) : posterior_(p), h_(h) {}
  const UnnormalisedPosterior& posterior_;
  thy_slq_t h_ = nullptr;
  thy_slq_stats stats_{};
};

}  // namespace thylacine
