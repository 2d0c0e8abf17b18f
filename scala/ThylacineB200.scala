/*
 * ThylacineB200.scala -- the reference-side binding of libthylacine_b200.so (Panama / java.lang.foreign, JDK 22+).
 *
 * Where it goes in gvonness/thylacine:  src/main/scala/thylacine/native/ThylacineB200.scala
 * What it binds:                        include/thylacine_b200.h of this repository, symbol for symbol.
 *
 * NOT COMPILED HERE: the build image has no JVM, scalac or sbt (SURVEY.md Appendix A).  The same C ABI is exercised through
 * ctypes (thylacine_b200/_capi.py) and a C++17 header (include/thylacine_b200.hpp); tests/test_abi.py checks that every
 * symbol bound below is exported by the library.  File:line citations are relative to the reference's
 * src/main/scala/thylacine/.
 *
 * Threading model: one JVM can drive the whole box.  Use ONE JVM THREAD PER GPU, bind it with setDevice(i) before it
 * builds its model / sampler, and hand every thread its communicator from commCreateAll (ncclCommInitAll underneath).
 * All calls block until the result is in the caller's off-heap buffer: wrap them in Async[F].blocking.
 */
package ai.entrolution.thylacine.native

import java.lang.foreign.*
import java.lang.foreign.ValueLayout.*
import java.lang.invoke.MethodHandle

private[thylacine] object ThylacineB200 {
  private val linker = Linker.nativeLinker()
  private val lib    = SymbolLookup.libraryLookup(sys.props.getOrElse("thylacine.b200.lib", "libthylacine_b200.so"), Arena.global())
  private def fn(name: String, res: MemoryLayout, args: MemoryLayout*): MethodHandle =
    linker.downcallHandle(lib.find(name).orElseThrow(), FunctionDescriptor.of(res, args*))

  private val I = JAVA_INT; private val L = JAVA_LONG; private val D = JAVA_DOUBLE; private val P = ADDRESS

  // ---- library / devices --------------------------------------------------------------------------------------------
  val version            = fn("thy_version", P)
  val lastError          = fn("thy_last_error", P)
  val kernelLaunchCount  = fn("thy_kernel_launch_count", L)
  val deviceCount        = fn("thy_device_count", I, P)
  val setDevice          = fn("thy_set_device", I, I)

  // ---- model construction (the case-class constructors) ----------------------------------------------------------------
  val modelCreate        = fn("thy_model_create", I, P)
  val modelDestroy       = fn("thy_model_destroy", I, P)
  val addGaussianPriorCi = fn("thy_model_add_gaussian_prior_ci", I, P, P, I, P, P)      // GaussianPrior.scala:62-75
  val addGaussianPriorCov= fn("thy_model_add_gaussian_prior_cov", I, P, P, I, P, P)     // GaussianPrior.scala:78-93
  val addUniformPrior    = fn("thy_model_add_uniform_prior", I, P, P, I, P, P)          // UniformPrior.scala:68-77 (max, min)
  val addCauchyPrior     = fn("thy_model_add_cauchy_prior", I, P, P, I, P, P)           // CauchyPrior.scala:50-63
  val addLikelihood      = fn("thy_model_add_likelihood", I, P, P)                      // likelihood/*.scala
  val registerLink       = fn("thy_model_register_link", I, P, P, P, P)                 // NonLinearForwardModel.scala:44-85 (closure -> CUDA source)
  val setOption          = fn("thy_model_set_option", I, P, I, I)
  val finalizeModel      = fn("thy_model_finalize", I, P)                               // Posterior.scala:40-48 label sort
  val collapse           = fn("thy_model_collapse", I, P, P)
  val gaussianPosterior  = fn("thy_model_gaussian_posterior", I, P, P, P, P, P, P)      // GaussianAnalyticPosterior.scala:141-169
  val dimension          = fn("thy_model_dimension", I, P, P)                           // Posterior.scala:37-38
  val labelLayout        = fn("thy_model_label_layout", I, P, P, P, P)                  // ModelParameterContext.scala:42-80
  val setStream          = fn("thy_model_set_stream", I, P, P)
  val synchronize        = fn("thy_model_synchronize", I, P)
  val modelLastError     = fn("thy_model_last_error", I, P, P, L)
  val lastKernelPath     = fn("thy_model_last_kernel_path", P, P)

  // ---- evaluation: ModelParameterPdf.logPdfAt / logPdfGradientAt (ModelParameterPdf.scala:33-46,84-95) -----------------
  val logPdfBatch        = fn("thy_logpdf_batch", I, P, L, P, P)
  val logPdfGradBatch    = fn("thy_logpdf_grad_batch", I, P, L, P, P, P)
  val logPdfGradAsync    = fn("thy_logpdf_grad_batch_async", I, P, L, P, P, P, P)       // two calls in flight: see waitFor
  val waitFor            = fn("thy_model_wait", I, P, L)
  val logPdfGradBatchDev = fn("thy_logpdf_grad_batch_dev", I, P, L, P, P, P)
  val samplePriors       = fn("thy_sample_priors", I, P, L, L, P)                       // Posterior.scala:64-65
  val samplePriorsDev    = fn("thy_sample_priors_dev", I, P, L, L, P)
  val enableTiming       = fn("thy_model_enable_timing", I, P, I)
  val kernelTime         = fn("thy_model_kernel_time", I, P, P, P)

  // ---- HmcmcSampledPosterior (HmcmcSampledPosterior.scala:63-75, HmcmcEngine.scala:55-214) -----------------------------
  val hmcConfig: StructLayout = MemoryLayout.structLayout(                                // HmcmcConfig.scala:20-25
    I.withName("stepsBetweenSamples"), I.withName("stepsInDynamicsSimulation"), I.withName("warmupStepCount"),
    MemoryLayout.paddingLayout(4), D.withName("dynamicsSimulationStepSize"))
  val hmcCreate          = fn("thy_hmc_create", I, P, P, L, L, P, P)
  val hmcDestroy         = fn("thy_hmc_destroy", I, P)
  val hmcSetUseGraph     = fn("thy_hmc_set_use_graph", I, P, I)
  val hmcSetFusedSteps   = fn("thy_hmc_set_fused_steps", I, P, I)
  val hmcBurnIn          = fn("thy_hmc_burn_in", I, P)
  val hmcRunTrajectories = fn("thy_hmc_run_trajectories", I, P, I)
  val hmcSample          = fn("thy_hmc_sample", I, P, I, P, I)                          // ModelParameterSampler.scala:36-37
  val hmcSampleDev       = fn("thy_hmc_sample_dev", I, P, I, P, I)
  val hmcStats           = fn("thy_hmc_stats", I, P, P, P, P)                           // HmcmcTelemetryUpdate.scala:20-25
  val hmcGetState        = fn("thy_hmc_get_state", I, P, P, P)
  val hmcSetState        = fn("thy_hmc_set_state", I, P, P, L)                          // checkpoint / resume
  val hmcTrajectoryCount = fn("thy_hmc_get_trajectory_counter", I, P, P)

  // ---- LeapfrogMcmcSampledPosterior (LeapfrogMcmcSampledPosterior.scala:72-106, LeapfrogMcmcEngine.scala:88-294) ------
  val lfmConfig: StructLayout = MemoryLayout.structLayout(                                // LeapfrogMcmcConfig.scala:20-24
    I.withName("stepsBetweenSamples"), I.withName("warmupStepCount"), I.withName("samplePoolSize"))
  val lfmCreate          = fn("thy_lfm_create", I, P, P, I, L, P, P)
  val lfmDestroy         = fn("thy_lfm_destroy", I, P)
  val lfmRunSweeps       = fn("thy_lfm_run_sweeps", I, P, I)
  val lfmSample          = fn("thy_lfm_sample", I, P, I, P)
  val lfmStats           = fn("thy_lfm_stats", I, P, P, P)
  val lfmGetPool         = fn("thy_lfm_get_pool", I, P, P, P)
  val lfmSetPool         = fn("thy_lfm_set_pool", I, P, P)

  // ---- multi-GPU plumbing (no counterpart in the reference: single JVM process, SURVEY 2a) --------------------------------
  val commUniqueId       = fn("thy_comm_unique_id", I, P)
  val commCreate         = fn("thy_comm_create", I, I, I, P, P)                         // one process per GPU
  val commCreateAll      = fn("thy_comm_create_all", I, I, P, P)                        // one JVM, one thread per GPU
  val commDestroy        = fn("thy_comm_destroy", I, P)
  val commRank           = fn("thy_comm_rank", I, P, P, P)
  val commAllReduceSum   = fn("thy_comm_all_reduce_sum", I, P, P, L, P)
  val momentsReset       = fn("thy_moments_reset_dev", I, P, P)
  val momentsAccumulate  = fn("thy_moments_accumulate_dev", I, P, L, P, P)
  val momentsFinish      = fn("thy_moments_finish", I, P, P, P, P, P, P)

  // ---- SlqIntegratedPosterior (SlqIntegratedPosterior.scala:76-125, SlqEngine.scala:107-469) ---------------------------
  val slqConfig: StructLayout = MemoryLayout.structLayout(                                // SlqConfig.scala:20-28 + extensions
    I.withName("poolSize"), I.withName("abscissaNumber"), D.withName("domainScalingIncrement"),
    D.withName("targetAcceptanceProbability"), I.withName("sampleParallelism"), I.withName("maxIterationCount"),
    I.withName("minIterationCount"), I.withName("constrainedWalkSteps"), I.withName("keepRetiredPoints"),
    I.withName("disableFusedWalk"), I.withName("referenceProposals"), MemoryLayout.paddingLayout(4))
  val slqStats: StructLayout = MemoryLayout.structLayout(                                 // SlqTelemetryUpdate.scala:20-29
    L.withName("iterations"), L.withName("total_retired"), L.withName("rounds"), L.withName("likelihood_evaluations"),
    D.withName("log_evidence_mean"), D.withName("log_evidence_std"), D.withName("neg_entropy_mean"), D.withName("min_logpdf"),
    D.withName("walk_scale"), D.withName("walk_acceptance"))
  val slqCreate          = fn("thy_slq_create", I, P, P, L, P, I, P, P)
  val slqDestroy         = fn("thy_slq_destroy", I, P)
  val slqRun             = fn("thy_slq_run", I, P, P)                                   // rebuildSampleSimulation, :76-80
  val slqRunRounds       = fn("thy_slq_run_rounds", I, P, I, P)
  val slqLivePoints      = fn("thy_slq_live_points", I, P, P, P, P)                     // SlqEngine.scala:120-143
  val slqRetired         = fn("thy_slq_retired", I, P, L, P, P)
  val slqRetiredLogX     = fn("thy_slq_retired_log_x", I, P, I, L, P, P)                // QuadratureAbscissa.scala:33-73
  val slqAbscissaResults = fn("thy_slq_abscissa_results", I, P, P, P)                   // QuadratureIntegrator.scala:29-62
  val slqIntegrate       = fn("thy_slq_integrate", I, P, P, P, D, P, P, P)              // SlqEngine.scala:460-466
  val slqSample          = fn("thy_slq_sample", I, P, I, L, P)                          // SamplingSimulation.scala:41-101
  val slqLoadBalance     = fn("thy_slq_load_balance", I, P, P, P)
  val slqSetTrace        = fn("thy_slq_set_trace", I, P, I)
  val slqPhaseTimes      = fn("thy_slq_phase_times", I, P, P, P)

  // ---- CartesianSurface (visualisation/CartesianSurface.scala:30-200) ------------------------------------------------------
  val surfaceCreate      = fn("thy_surface_create", I, I, P, I, P, P)
  val surfaceDestroy     = fn("thy_surface_destroy", I, P)
  val surfaceAddSamples  = fn("thy_surface_add_samples", I, P, I, P, L, P, D, D)
  val surfaceAddSamplesDev = fn("thy_surface_add_samples_dev", I, P, I, P, L, P, D, D)
  val surfaceResults     = fn("thy_surface_results", I, P, I, P)

  // struct thy_likelihood_desc (header): 5 ints, 4 bytes padding, 6 pointers, 1 double
  val likelihoodDesc: StructLayout = MemoryLayout.structLayout(
    I.withName("distribution"), I.withName("link"), I.withName("jacobian"), I.withName("n"), I.withName("nblocks"),
    MemoryLayout.paddingLayout(4), P.withName("labels"), P.withName("dims"), P.withName("blocks"), P.withName("offset"),
    P.withName("measurements"), P.withName("uncertainties"), D.withName("differential"))

  // thy_slq_integrand: double (*)(double scaledLikelihood, void* user) -- upcall stub for a Scala Double => Double
  private val integrandDesc = FunctionDescriptor.of(D, D, P)
  final class Integrand(f: Double => Double) { def apply(v: Double, user: MemorySegment): Double = f(v) }
  def integrandStub(f: Double => Double, arena: Arena): MemorySegment = {
    val mh = java.lang.invoke.MethodHandles.lookup().bind(new Integrand(f), "apply",
      java.lang.invoke.MethodType.methodType(classOf[Double], classOf[Double], classOf[MemorySegment]))
    linker.upcallStub(mh, integrandDesc, arena)
  }

  /** What the reference asserts comes back as status 1 (AssertionError), a missing label as 2 (RuntimeException,
    * IndexedCollection.scala:25-31); everything else is an IllegalStateException carrying the library's message. */
  def check(status: Int): Unit = if (status != 0) {
    val msg = lastError.invoke().asInstanceOf[MemorySegment].reinterpret(4096).getString(0)
    status match {
      case 1 => throw new AssertionError(msg)
      case 2 => throw new RuntimeException(msg)
      case 7 => throw new UnsupportedOperationException(s"no CUDA device: thylacine_b200 has no CPU fallback ($msg)")
      case _ => throw new IllegalStateException(s"thylacine_b200 status $status: $msg")
    }
  }

  def doubles(arena: Arena, xs: Iterable[Double]): MemorySegment = {
    val a   = xs.toArray
    val seg = arena.allocate(JAVA_DOUBLE, a.length.toLong)
    MemorySegment.copy(a, 0, seg, JAVA_DOUBLE, 0L, a.length)
    seg
  }
}

/** A posterior whose evaluation runs on the GPU.  Mixes into the reference where UnnormalisedPosterior does
  * (posterior/UnnormalisedPosterior.scala:28-47); `priors` / `likelihoods` are the reference's own case classes and are
  * only read at construction.  The single-point members keep the reference's signatures; batched callers (HMC chains, pool
  * members, live points) use logPdfGradBatch. */
final class B200Posterior(priorSpecs: Seq[B200Posterior.PriorSpec], likelihoodSpecs: Seq[B200Posterior.LikelihoodSpec]) extends AutoCloseable {
  import ThylacineB200.*
  private val arena = Arena.ofShared()
  val handle: MemorySegment = {
    val out = arena.allocate(ADDRESS)
    check(modelCreate.invoke(out).asInstanceOf[Int])
    val m = out.get(ADDRESS, 0)
    priorSpecs.foreach {
      case B200Posterior.GaussianCi(label, values, cis) =>
        check(addGaussianPriorCi.invoke(m, arena.allocateFrom(label), values.size, doubles(arena, values), doubles(arena, cis)).asInstanceOf[Int])
      case B200Posterior.Uniform(label, maxBounds, minBounds) =>
        check(addUniformPrior.invoke(m, arena.allocateFrom(label), maxBounds.size, doubles(arena, maxBounds), doubles(arena, minBounds)).asInstanceOf[Int])
      case B200Posterior.Cauchy(label, values, cis) =>
        check(addCauchyPrior.invoke(m, arena.allocateFrom(label), values.size, doubles(arena, values), doubles(arena, cis)).asInstanceOf[Int])
    }
    likelihoodSpecs.foreach { l =>
      val desc = arena.allocate(likelihoodDesc)
      val nb   = l.blocks.size
      val labels = arena.allocate(ADDRESS, nb.toLong); val dims = arena.allocate(JAVA_INT, nb.toLong); val blocks = arena.allocate(ADDRESS, nb.toLong)
      l.blocks.zipWithIndex.foreach { case ((label, width, rowMajor), i) =>
        labels.setAtIndex(ADDRESS, i.toLong, arena.allocateFrom(label)); dims.setAtIndex(JAVA_INT, i.toLong, width)
        blocks.setAtIndex(ADDRESS, i.toLong, doubles(arena, rowMajor))
      }
      desc.set(JAVA_INT, 0, l.distribution); desc.set(JAVA_INT, 4, l.link); desc.set(JAVA_INT, 8, l.jacobian)
      desc.set(JAVA_INT, 12, l.measurements.size); desc.set(JAVA_INT, 16, nb)
      desc.set(ADDRESS, 24, labels); desc.set(ADDRESS, 32, dims); desc.set(ADDRESS, 40, blocks)
      desc.set(ADDRESS, 48, l.offset.map(doubles(arena, _)).getOrElse(MemorySegment.NULL))
      desc.set(ADDRESS, 56, doubles(arena, l.measurements)); desc.set(ADDRESS, 64, doubles(arena, l.uncertainties))
      desc.set(JAVA_DOUBLE, 72, l.differential)
      check(addLikelihood.invoke(m, desc).asInstanceOf[Int])
    }
    check(finalizeModel.invoke(m).asInstanceOf[Int])
    m
  }
  val domainDimension: Int = { val o = arena.allocate(JAVA_INT); check(dimension.invoke(handle, o).asInstanceOf[Int]); o.get(JAVA_INT, 0) }

  /** (offset, width) of a label in the flat, label-sorted vector: ModelParameterContext.scala:42-80. */
  def layout(label: String): (Int, Int) = {
    val off = arena.allocate(JAVA_INT); val dim = arena.allocate(JAVA_INT)
    check(labelLayout.invoke(handle, arena.allocateFrom(label), off, dim).asInstanceOf[Int]); (off.get(JAVA_INT, 0), dim.get(JAVA_INT, 0))
  }

  /** Posterior.logPdfAt / logPdfGradientAt (Posterior.scala:50-73) for B points at once, row-major X(b * d + j). */
  def logPdfGradBatch(x: Array[Double]): (Array[Double], Array[Double]) = {
    val b = x.length / domainDimension
    val call = Arena.ofConfined()
    try {
      val xs = call.allocate(JAVA_DOUBLE, x.length.toLong); MemorySegment.copy(x, 0, xs, JAVA_DOUBLE, 0L, x.length)
      val lp = call.allocate(JAVA_DOUBLE, b.toLong); val g = call.allocate(JAVA_DOUBLE, x.length.toLong)
      check(ThylacineB200.logPdfGradBatch.invoke(handle, b.toLong, xs, lp, g).asInstanceOf[Int])
      (lp.toArray(JAVA_DOUBLE), g.toArray(JAVA_DOUBLE))
    } finally call.close()
  }
  override def close(): Unit = { modelDestroy.invoke(handle); arena.close() }
}

object B200Posterior {
  sealed trait PriorSpec
  final case class GaussianCi(label: String, values: Vector[Double], confidenceIntervals: Vector[Double]) extends PriorSpec
  final case class Uniform(label: String, maxBounds: Vector[Double], minBounds: Vector[Double]) extends PriorSpec
  final case class Cauchy(label: String, values: Vector[Double], confidenceIntervals: Vector[Double]) extends PriorSpec
  /** blocks: (label, width, n x width coefficients row-major); distribution / link / jacobian: the header's enums. */
  final case class LikelihoodSpec(distribution: Int, link: Int, jacobian: Int, blocks: Seq[(String, Int, Vector[Double])],
                                  offset: Option[Vector[Double]], measurements: Vector[Double], uncertainties: Vector[Double],
                                  differential: Double = 0.0)
}
