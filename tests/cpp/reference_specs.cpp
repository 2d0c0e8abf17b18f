// The reference's own component specs, driven from compiled code through include/thylacine_b200.hpp (the C++ mirror of
// the reference's operator surface over the C ABI).  Expected values: /root/reference/src/test/scala/thylacine/model/
//   components/ComponentFixture.scala:34-111, prior/GaussianPriorSpec.scala:49-58,
//   likelihood/GaussianLinearLikelihoodSpec.scala:32-48, posterior/GaussianAnalyticPosteriorSpec.scala:34-50
// (also committed as tests/golden/reference_specs.json).
//   reference_specs            all specs on cuda:0
//   reference_specs --no-gpu   only checks that the product path fails loudly without a device (no CPU fallback)
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>

#include "thylacine_b200.hpp"

using namespace thylacine;

static int failures = 0;
#define EXPECT(cond)                                                              \
  do {                                                                            \
    if (!(cond)) {                                                                \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);      \
      ++failures;                                                                 \
    }                                                                             \
  } while (0)

static bool throws_with(thy_status want, const std::function<void()>& f) {
  try {
    f();
  } catch (const Error& e) {
    return e.status == want;
  }
  return false;
}

static Prior fooPrior() { return GaussianPrior::fromConfidenceIntervals("foo", {1, 2}, {3, 5}); }          // ComponentFixture:34-39
static Prior barPrior() { return GaussianPrior::fromConfidenceIntervals("bar", {5}, {0.1}); }              // :41-46
static Likelihood fooLikelihood() {                                                                        // :48-54
  return GaussianLinearLikelihood::of(Matrix({{1, 3}, {2, 4}}), {7, 10}, {0.01, 0.01}, "foo");
}
static Likelihood barLikelihood() {                                                                        // :80-86
  return GaussianLinearLikelihood::of(Matrix({{3}, {4}}), {15, 20}, {0.00001, 0.00001}, "bar");
}

int main(int argc, char** argv) {
  if (argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0) {
    // no CPU fallback: building a posterior needs the device
    EXPECT(throws_with(THY_ERR_NO_DEVICE, [] { UnnormalisedPosterior p({fooPrior()}, {fooLikelihood()}); }));
    // argument checks happen before the device is touched: what the reference asserts is still reported
    EXPECT(throws_with(THY_ERR_INVALID_ARGUMENT, [] {
      UnnormalisedPosterior p({GaussianPrior::fromConfidenceIntervals("foo", {1, 2}, {3, -5})});     // RecordedData.scala:57
    }));
    EXPECT(throws_with(THY_ERR_INVALID_ARGUMENT, [] { UnnormalisedPosterior p({fooPrior(), fooPrior()}); }));
    std::printf(failures ? "cpp api FAILED\n" : "cpp api ok (no gpu)\n");
    return failures ? 1 : 0;
  }

  {  // GaussianPriorSpec.scala:49-58 -- exact equality, as the reference asserts
    UnnormalisedPosterior post({GaussianPrior::fromConfidenceIntervals("p", {2, 3, 4}, {4, 3, 2})});
    const Vector g = post.logPdfGradientAt({{"p", {3, 2, 5}}}).at("p");
    EXPECT(g[0] == -0.25 && g[1] == 4.0 / 9.0 && g[2] == -1.0);
  }
  {  // GaussianLinearLikelihoodSpec.scala:32-48 through a zero-gradient uniform prior (UniformDistribution.scala:78-81)
    UnnormalisedPosterior post({UniformPrior::fromBounds("foo", {5, 5}, {-5, -5})}, {fooLikelihood()});
    const Vector g0 = post.logPdfGradientAt({{"foo", {1, 2}}}).at("foo");
    const Vector g1 = post.logPdfGradientAt({{"foo", {3, 2}}}).at("foo");
    EXPECT(g0[0] == 0.0 && g0[1] == 0.0);
    EXPECT(g1[0] == -4e5 && g1[1] == -8.8e5);
    const double lv = -2.0 * std::log(10.0);
    EXPECT(std::fabs(post.logPdfAt({{"foo", {1, 2}}}) - (8.758757666686726 + lv)) < 1e-12);
    EXPECT(std::fabs(post.logPdfAt({{"foo", {3, 2}}}) - (-399991.2412423333 + lv)) < 1e-7);
    EXPECT(post.logPdfAt({{"foo", {5.0, 0}}}) == -INFINITY);   // half-open box [lo, hi) (UniformDistribution.scala:63)
  }
  {  // label-sorted flattening and the posterior sum (Posterior.scala:40-73)
    UnnormalisedPosterior post({fooPrior(), barPrior()}, {fooLikelihood(), barLikelihood()});
    EXPECT(post.domainDimension() == 3);
    EXPECT(post.layout().at("bar") == std::make_pair(0, 1) && post.layout().at("foo") == std::make_pair(1, 2));
    const ModelParameters g = post.logPdfGradientAt({{"foo", {1, 2}}, {"bar", {5}}});
    EXPECT(g.at("foo")[0] == 0.0 && g.at("foo")[1] == 0.0 && g.at("bar")[0] == 0.0);
    EXPECT(throws_with(THY_ERR_UNKNOWN_LABEL, [&] { post.logPdfAt({{"baz", {1}}}); }));
  }
  {  // GaussianAnalyticPosteriorSpec.scala:34-50: mean within 0.1, covariance within 0.01
    GaussianAnalyticPosterior post({fooPrior(), barPrior()}, {fooLikelihood(), barLikelihood()});
    const ModelParameters m = post.mean();
    EXPECT(std::fabs(m.at("foo")[0] - 1) < 0.1 && std::fabs(m.at("foo")[1] - 2) < 0.1 && std::fabs(m.at("bar")[0] - 5) < 0.1);
    for (double c : post.covarianceStridedVector()) EXPECT(std::fabs(c) < 0.01);
  }
  {  // what the reference asserts
    EXPECT(throws_with(THY_ERR_INVALID_ARGUMENT, [] { UnnormalisedPosterior p({UniformPrior::fromBounds("u", {0}, {1})}); }));
    EXPECT(throws_with(THY_ERR_UNKNOWN_LABEL, [] {
      UnnormalisedPosterior p({barPrior()}, {fooLikelihood()});   // likelihood reads a label no prior owns
    }));
    EXPECT(throws_with(THY_ERR_INVALID_ARGUMENT, [] {
      GaussianLinearLikelihood::of(Matrix({{1, 3}, {2, 4}}), {7, 10, 11}, {0.01, 0.01, 0.01}, "foo");   // :43-45
    }));
  }
  {  // the collapsed posterior is the same function; HMC and SLQ run on either
    const int d = 6, n = 40;
    Vector A((size_t)n * d), y(n), ci(n, 0.4);
    uint64_t s = 12345;
    auto rnd = [&] { s = s * 6364136223846793005ull + 1442695040888963407ull; return ((double)(s >> 11) / 9007199254740992.0) * 2 - 1; };
    for (auto& a : A) a = rnd();
    for (auto& v : y) v = rnd();
    std::vector<Prior> priors{GaussianPrior::fromConfidenceIntervals("theta", Vector(d, 0.0), Vector(d, 6.0))};
    std::vector<Likelihood> liks{GaussianLinearLikelihood::of(Matrix(n, d, A), y, ci, "theta")};
    UnnormalisedPosterior post(priors, liks);
    UnnormalisedPosterior col = post.collapsed();
    Vector X((size_t)32 * d);
    for (auto& v : X) v = rnd();
    auto a = post.logPdfGradBatch(X), b = col.logPdfGradBatch(X);
    double el = 0, eg = 0, gs = 0;
    for (size_t i = 0; i < a.first.size(); ++i) el = std::fmax(el, std::fabs(a.first[i] - b.first[i]) / std::fmax(1.0, std::fabs(a.first[i])));
    for (size_t i = 0; i < a.second.size(); ++i) { eg = std::fmax(eg, std::fabs(a.second[i] - b.second[i])); gs = std::fmax(gs, std::fabs(a.second[i])); }
    EXPECT(el < 1e-10 && eg < 1e-10 * gs);

    GaussianAnalyticPosterior analytic(priors, liks);
    const Vector mean = analytic.mean().at("theta");
    double sd_min = 1e300;
    for (int j = 0; j < d; ++j) sd_min = std::fmin(sd_min, std::sqrt(analytic.covarianceStridedVector()[(size_t)j * d + j]));
    const ModelParameters start = analytic.mean();   // chains start at the mode; 60 warm-up trajectories decorrelate them
    HmcmcSampledPosterior hmc({1, 12, 60, 0.2 * sd_min}, post, 512, 7, &start);
    const Vector smp = hmc.sampleChains(2);
    Vector avg(d, 0.0);
    for (size_t r = 0; r < smp.size() / d; ++r)
      for (int j = 0; j < d; ++j) avg[j] += smp[r * d + j] / (double)(smp.size() / d);
    for (int j = 0; j < d; ++j) {
      const double sd = std::sqrt(analytic.covarianceStridedVector()[(size_t)j * d + j]);
      EXPECT(std::fabs(avg[j] - mean[j]) < 5.0 * sd / std::sqrt(512.0));
    }
    EXPECT(hmc.telemetry()[0].jumpAttempts > 0);

    SlqIntegratedPosterior slq = SlqIntegratedPosterior::of({4000, 6, 0.05, 0.3, 200, 40 * 4000, 4000}, col, 3, 30);
    slq.rebuildSampleSimulation();
    const double tol = 5.0 * std::sqrt(slq.stats().neg_entropy_mean / 4000.0) + 3.0 * slq.stats().log_evidence_std / std::sqrt(6.0);
    EXPECT(std::fabs(slq.logEvidence() - analytic.logEvidence()) < tol);

    LeapfrogMcmcSampledPosterior lfm = LeapfrogMcmcSampledPosterior::of({2, 20, 256}, THY_DISTANCE_EUCLIDEAN, post, 5);
    EXPECT(lfm.sample(4).size() == 4);
  }
  std::printf(failures ? "cpp api FAILED (%d)\n" : "cpp api ok\n", failures);
  return failures ? 1 : 0;
}
