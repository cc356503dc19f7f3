;;   B200-native device engines for Bayadera, over the C ABI of libbayadera_b200.so (include/bayadera_b200.h).
;;
;;   Drop-in twin of uncomplicate.bayadera.internal.device.nvidia-gtx: the same protocols
;;   (uncomplicate.bayadera.internal.protocols), the same constructor arities, the same result types, so that
;;   uncomplicate.bayadera.core / mcmc / library and the Midje tests run unchanged on top of it.  Where the reference
;;   compiles its .cu sources with ClojureCUDA and launches kernels itself, this namespace makes ONE call per
;;   protocol method into the shared library, which compiles the model's C sources with NVRTC for sm_100a and owns
;;   every kernel (DESIGN.md).  Reference line numbers below are those of
;;   src/clojure/uncomplicate/bayadera/internal/device/nvidia_gtx.clj ("G/") of uncomplicate/bayadera 0.4.0-SNAPSHOT.
;;
;;   Requirements on the classpath: the reference's own dependencies (project.clj) plus net.java.dev.jna/jna;
;;   libbayadera_b200.so on jna.library.path.  NOT load-tested in the build container (it has no JVM): every call
;;   made here is exercised through the identical ctypes calls of the Python mirror (bayadera_b200/engine.py) in
;;   tests/, including inside a non-primary CUDA context with borrowed device buffers (tests/test_gpu_context.py),
;;   which is exactly the situation ClojureCUDA's with-default creates.

(ns ^{:author "bayadera_b200"}
    uncomplicate.bayadera.internal.device.b200
  (:require [uncomplicate.commons
             [core :refer [Releaseable release with-release let-release Info info]]
             [utils :refer [dragan-says-ex]]]
            [uncomplicate.clojurecuda.core :refer [in-context ctx-device multiprocessor-count max-block-dim-x]]
            [uncomplicate.neanderthal.internal.api :as na]
            [uncomplicate.neanderthal
             [core :refer [vctr ge ncols mrows dim transfer transfer! entry!]]
             [native :refer [fv fge]]
             [block :refer [buffer offset stride]]
             [cuda :refer [cuda-float]]]
            [uncomplicate.bayadera.internal.protocols :refer :all])
  (:import [com.sun.jna Library Native Pointer]
           [com.sun.jna.ptr PointerByReference DoubleByReference LongByReference IntByReference]))

;; ============================ The C ABI ======================================
;; One method per entry point of include/bayadera_b200.h (same names, same argument order).  Device pointers
;; (CUdeviceptr of the CURRENT ClojureCUDA context) travel as long / Pointer; host arrays as float[] / int[].

(definterface Bay
  (^String bay_last_error [])
  (^String bay_version [])
  ;; engine
  (^int bay_engine_create [^int device ^long stream ^int wgs ^com.sun.jna.ptr.PointerByReference out])
  (^int bay_engine_create_current [^long stream ^int wgs ^com.sun.jna.ptr.PointerByReference out])
  (^int bay_engine_release [^com.sun.jna.Pointer e])
  (^int bay_engine_processing_elements [^com.sun.jna.Pointer e ^com.sun.jna.ptr.LongByReference out])
  (^int bay_engine_stream [^com.sun.jna.Pointer e ^com.sun.jna.ptr.LongByReference out])
  (^int bay_engine_synchronize [^com.sun.jna.Pointer e])
  (^int bay_nccl_unique_id [^bytes id])
  (^int bay_engine_comm_init [^com.sun.jna.Pointer e ^bytes id ^int nranks ^int rank])
  ;; model
  (^int bay_model_compile [^com.sun.jna.Pointer e ^"[Ljava.lang.String;" srcs ^int nsrc ^String logfn-name
                           ^int dim ^int params-size ^int flags ^com.sun.jna.ptr.PointerByReference out])
  (^int bay_model_release [^com.sun.jna.Pointer m])
  (^int bay_model_uses_quadform [^com.sun.jna.Pointer m])
  ;; sampler
  (^int bay_sampler_create [^com.sun.jna.Pointer m ^int seed ^long walkers ^floats params ^long n
                            ^com.sun.jna.ptr.PointerByReference out])
  (^int bay_sampler_create_dev [^com.sun.jna.Pointer m ^int seed ^long walkers ^long params-dev ^long n
                                ^com.sun.jna.ptr.PointerByReference out])
  (^int bay_sampler_release [^com.sun.jna.Pointer s])
  (^int bay_init [^com.sun.jna.Pointer s ^int seed])
  (^int bay_init_position_uniform [^com.sun.jna.Pointer s ^int seed ^floats limits])
  (^int bay_init_position_from [^com.sun.jna.Pointer s ^com.sun.jna.Pointer other])
  (^int bay_burn_in [^com.sun.jna.Pointer s ^long n ^float a])
  (^int bay_anneal [^com.sun.jna.Pointer s ^floats temperature ^long n ^float a])
  (^int bay_acc_rate [^com.sun.jna.Pointer s ^float a ^com.sun.jna.ptr.DoubleByReference out])
  (^int bay_run_sampler [^com.sun.jna.Pointer s ^long n ^float a ^com.sun.jna.ptr.DoubleByReference acc
                         ^floats tau ^floats mean ^floats sigma ^com.sun.jna.ptr.LongByReference lag])
  (^int bay_last_means [^com.sun.jna.Pointer s ^floats means ^long n])
  (^int bay_init_move [^com.sun.jna.Pointer s ^float a])
  (^int bay_move [^com.sun.jna.Pointer s])
  (^int bay_move_bare [^com.sun.jna.Pointer s])
  (^int bay_set_temperature [^com.sun.jna.Pointer s ^float t])
  (^int bay_sample [^com.sun.jna.Pointer s ^long n ^com.sun.jna.Pointer out ^int out-is-device])
  (^int bay_histogram [^com.sun.jna.Pointer s ^long cycles ^floats limits ^floats pdf ^floats ranks])
  (^int bay_mean [^com.sun.jna.Pointer s ^floats out])
  (^int bay_variance [^com.sun.jna.Pointer s ^floats out])
  (^int bay_sd [^com.sun.jna.Pointer s ^floats out])
  (^int bay_info [^com.sun.jna.Pointer s ^com.sun.jna.ptr.LongByReference walkers
                  ^com.sun.jna.ptr.LongByReference iterations])
  (^int bay_mix [^com.sun.jna.Pointer s ^long step ^double dimension-power ^int schedule ^double schedule-power
                 ^double a ^double min-acc ^double max-acc ^com.sun.jna.ptr.DoubleByReference out-a
                 ^com.sun.jna.ptr.DoubleByReference out-acc ^com.sun.jna.ptr.DoubleByReference out-acc-2])
  (^int bay_hdi [^com.sun.jna.Pointer s ^double mass ^ints counts ^ints nregions ^floats regions ^int max-regions])
  ;; dataset / acor engines (data-is-device = 1: the pointer is a CUdeviceptr of the current context)
  (^int bay_dataset_mean [^com.sun.jna.Pointer e ^com.sun.jna.Pointer data ^int data-is-device ^long m ^long n
                          ^long offset ^long ld ^floats out])
  (^int bay_dataset_variance [^com.sun.jna.Pointer e ^com.sun.jna.Pointer data ^int data-is-device ^long m ^long n
                              ^long offset ^long ld ^floats out])
  (^int bay_dataset_histogram [^com.sun.jna.Pointer e ^com.sun.jna.Pointer data ^int data-is-device ^long m ^long n
                               ^long offset ^long ld ^floats limits ^floats pdf ^floats ranks ^ints counts])
  (^int bay_acor [^com.sun.jna.Pointer e ^floats series ^long dim ^long n ^floats tau ^floats mean ^floats sigma
                  ^com.sun.jna.ptr.LongByReference lag])
  ;; density / likelihood / direct-sampler engines
  (^int bay_model_logfn [^com.sun.jna.Pointer m ^floats params ^long nparams ^floats x ^long n ^floats out])
  (^int bay_model_density [^com.sun.jna.Pointer m ^floats params ^long nparams ^floats x ^long n ^int exponentiate
                           ^floats out])
  (^int bay_model_evidence [^com.sun.jna.Pointer m ^floats params ^long nparams ^floats x ^long n
                            ^com.sun.jna.ptr.DoubleByReference out])
  (^int bay_model_density_dev [^com.sun.jna.Pointer m ^long params-dev ^long nparams ^long x-dev ^long n
                               ^int exponentiate ^long out-dev])
  (^int bay_model_evidence_dev [^com.sun.jna.Pointer m ^long params-dev ^long nparams ^long x-dev ^long n
                                ^com.sun.jna.ptr.DoubleByReference out])
  (^int bay_direct_sample [^com.sun.jna.Pointer e ^int family ^int seed ^floats params ^int nparams ^long n
                           ^com.sun.jna.Pointer out ^int out-is-device]))

(def ^Bay bay (Native/load "bayadera_b200" Bay))

;; status codes of bay_status
(def ^:const BAY_EINVAL_WALKERS -2)
(def ^:const BAY_EACOR_TOO_SHORT -3)

;; model flags (extra keys of a model's args map, ignored by the reference; SURVEY Appendix C)
(def ^:const FAST-MATH 0x1)
(def ^:const ROW-ADDITIVE 0x2)
(def ^:const GLM-LOGISTIC 0x4)
(def ^:const GLM-POISSON 0x8)
(def ^:const QUADFORM 0x10)

(defmacro ^:private ok!
  "Status code -> the exception types the reference throws (G/:275-278, 609-610; ex-info elsewhere)."
  [call]
  `(let [rc# (int ~call)]
     (when-not (zero? rc#)
       (let [msg# (.bay_last_error bay)]
         (if (or (= rc# BAY_EINVAL_WALKERS) (= rc# BAY_EACOR_TOO_SHORT))
           (throw (IllegalArgumentException. ^String msg#))
           (throw (ex-info msg# {:code rc#})))))))

(defn ^:private device-pointer
  "The raw CUdeviceptr behind a ClojureCUDA buffer / Neanderthal cuda-float block, as a long."
  ^long [cu-buf]
  (long (uncomplicate.clojurecuda.internal.protocols/ptr cu-buf)))

(defn ^:private dev-addr
  "CUdeviceptr of the first entry of a cuda-float vector / dense matrix (buffer + offset)."
  ^long [x]
  (+ (device-pointer (buffer x)) (* Float/BYTES (long (offset x)))))

(defn ^:private dev-ptr ^Pointer [x]
  (Pointer. (dev-addr x)))

(defn ^:private host-floats
  "Column-major float[] copy of a (host or device) Neanderthal vector / matrix."
  ^floats [x]
  (with-release [h (transfer x)]
    (float-array (seq h))))

(defn ^:private model-flags ^long [model]
  ;; DeviceDistributionModel keeps the args map it was built from; models that want the row-additive
  ;; tensor-core paths carry :flags there (GLM-LOGISTIC / GLM-POISSON / QUADFORM).  The reference compiles with
  ;; -use_fast_math (G/:630-633), hence FAST-MATH always.
  (bit-or FAST-MATH (long (or (:flags model) 0))))

(defn ^:private compile-model
  "gtx-stretch-factory's NVRTC step (G/:747-757): model sources first, the engine's kernels after."
  ^Pointer [^Pointer eng model logfn-name]
  (let [out (PointerByReference.)
        srcs (into-array String (source model))]
    (ok! (.bay_model_compile bay eng srcs (alength srcs) (str logfn-name)
                             (int (dimension model)) (int (params-size model)) (int (model-flags model)) out))
    (.getValue out)))

;; ============================ Direct sampler (G/:48-63) ======================

(def ^:private direct-families {"uniform" 0 "gaussian" 1 "exponential" 2 "erlang" 3})

(deftype B200DirectSamplerEngine [ctx ^Pointer eng ^long family]
  Releaseable
  (release [_] true)
  RandomSamplerEngine
  (sample [this seed cu-params res]
    (if (and (= 0 (rem (ncols res) 4)) (= 0 (rem (offset res) 4)))
      (in-context
       ctx
       (let [p (host-floats cu-params)]
         (ok! (.bay_direct_sample bay eng (int family) (int seed) p (alength p) (long (ncols res)) (dev-ptr res) 1))
         res))
      (dragan-says-ex "GTX direct sampler supports only matrices with ncols and offset that are multiple of 4."
                      {:ncols (ncols res) :offset (offset res)}))))

;; ============================ Distribution engine (G/:67-104) ================

(deftype B200DistributionEngine [ctx ^Pointer modl dist-model]
  Releaseable
  (release [_] (ok! (.bay_model_release bay modl)) true)
  ModelProvider
  (model [_] dist-model)
  DensityEngine
  ;; params, the DIM x n point matrix (stride = DIM, as the reference's kernels assume, G/:83-104) and the result are
  ;; cuda-float blocks of ctx: they are handed across as device addresses, nothing is staged through the host
  (log-density [this cu-params x]
    (in-context
     ctx
     (let-release [res (vctr cu-params (ncols x))]
       (ok! (.bay_model_density_dev bay modl (dev-addr cu-params) (long (dim cu-params)) (dev-addr x)
                                    (long (ncols x)) 0 (dev-addr res)))
       res)))
  (density [this cu-params x]
    (in-context
     ctx
     (let-release [res (vctr cu-params (ncols x))]
       (ok! (.bay_model_density_dev bay modl (dev-addr cu-params) (long (dim cu-params)) (dev-addr x)
                                    (long (ncols x)) 1 (dev-addr res)))
       res))))

;; ============================ Likelihood engine (G/:106-141) =================
;; The model handle is compiled with a LOGFN that ignores the hyper-parameters and calls the model's loglik on
;; params[0, data-len): (likelihood-source model) below; log-density = loglik, density = lik, evidence = mean lik.

(defn ^:private likelihood-source [model]
  (format (str "extern \"C\" { inline REAL bay_loglik_engine(const uint32_t data_len, const uint32_t params_len, "
               "const REAL* params, const uint32_t dim, const REAL* x) { return %s(data_len, params, dim, x); } }\n")
          (loglik model)))

(deftype B200LikelihoodEngine [ctx ^Pointer modl dist-model]
  Releaseable
  (release [_] (ok! (.bay_model_release bay modl)) true)
  ModelProvider
  (model [_] dist-model)
  DensityEngine
  (log-density [this cu-data x]
    (in-context
     ctx
     (let-release [res (vctr cu-data (ncols x))]
       (ok! (.bay_model_density_dev bay modl (dev-addr cu-data) (long (dim cu-data)) (dev-addr x)
                                    (long (ncols x)) 0 (dev-addr res)))
       res)))
  (density [this cu-data x]
    (in-context
     ctx
     (let-release [res (vctr cu-data (ncols x))]
       (ok! (.bay_model_density_dev bay modl (dev-addr cu-data) (long (dim cu-data)) (dev-addr x)
                                    (long (ncols x)) 1 (dev-addr res)))
       res)))
  LikelihoodEngine
  (evidence [this cu-data x]
    (in-context
     ctx
     (let [out (DoubleByReference.)]
       (ok! (.bay_model_evidence_dev bay modl (dev-addr cu-data) (long (dim cu-data)) (dev-addr x)
                                     (long (ncols x)) out))
       (.getValue out)))))

;; ============================ Dataset engine (G/:145-228) ====================
;; data-matrix is a Neanderthal cuda-float ge of the current context: its buffer pointer, offset and stride go
;; straight across (bay_dataset_*: element (row, col) at data[offset + ld*col + row]).

(deftype B200DatasetEngine [ctx ^Pointer eng ^long WGS]
  Releaseable
  (release [_] true)
  DatasetEngine
  (data-mean [this data-matrix]
    (in-context
     ctx
     (let [m (mrows data-matrix) out (float-array m)]
       (ok! (.bay_dataset_mean bay eng (Pointer. (device-pointer (buffer data-matrix))) 1 (long m)
                               (long (ncols data-matrix)) (long (offset data-matrix)) (long (stride data-matrix)) out))
       (let-release [res (vctr data-matrix m)]
         (transfer! out res)))))
  (data-variance [this data-matrix]
    (in-context
     ctx
     (let [m (mrows data-matrix) out (float-array m)]
       (ok! (.bay_dataset_variance bay eng (Pointer. (device-pointer (buffer data-matrix))) 1 (long m)
                                   (long (ncols data-matrix)) (long (offset data-matrix)) (long (stride data-matrix)) out))
       (let-release [res (vctr data-matrix m)]
         (transfer! out res)))))
  EstimateEngine
  (histogram [this data-matrix]
    (in-context
     ctx
     (let [m (mrows data-matrix)
           lim (float-array (* 2 m)) pdf (float-array (* WGS m)) ranks (float-array (* WGS m))]
       (ok! (.bay_dataset_histogram bay eng (Pointer. (device-pointer (buffer data-matrix))) 1 (long m)
                                    (long (ncols data-matrix)) (long (offset data-matrix)) (long (stride data-matrix))
                                    lim pdf ranks nil))
       ;; host matrices, column = dimension, like (transfer limits) ... in G/:228
       (->Histogram (fge 2 m lim) (fge WGS m pdf) (fge WGS m ranks))))))

;; ============================ Acor engine (G/:230-278) =======================

(deftype B200AcorEngine [ctx ^Pointer eng]
  Releaseable
  (release [_] true)
  AcorEngine
  (acor [_ data-matrix]
    (in-context
     ctx
     (let [d (mrows data-matrix) n (ncols data-matrix)
           series (host-floats data-matrix)
           tau (float-array d) mean (float-array d) sigma (float-array d) lag (LongByReference.)]
       ;; BAY_EACOR_TOO_SHORT -> IllegalArgumentException with the reference's text (G/:275-278)
       (ok! (.bay_acor bay eng series (long d) (long n) tau mean sigma lag))
       (->Autocorrelation (fv tau) (fv mean) (fv sigma) n (.getValue lag))))))

;; ============================ Stretch sampler (G/:282-541) ===================

(deftype B200Stretch [ctx ^Pointer handle neanderthal-factory cu-model ^long DIM ^long WGS ^long walker-count]
  Releaseable
  (release [_]
    (in-context ctx (ok! (.bay_sampler_release bay handle)))
    true)
  Info
  (info [this]
    (let [w (LongByReference.) it (LongByReference.)]
      (ok! (.bay_info bay handle w it))
      {:walker-count (.getValue w)
       :iteration-counter (.getValue it)}))
  ModelProvider
  (model [this]
    cu-model)
  MCMCStretch
  ;; the reference hands in a means accumulator buffer (G/:340); the library keeps the per-step means itself
  ;; (bay_last_means), so the argument is accepted and ignored
  (init-move! [this cu-means-acc a]
    (in-context ctx (ok! (.bay_init_move bay handle (float a))))
    this)
  (move! [this]
    (in-context ctx (ok! (.bay_move bay handle)))
    this)
  (move-bare! [this]
    (in-context ctx (ok! (.bay_move_bare bay handle)))
    this)
  (set-temperature! [this t]
    (ok! (.bay_set_temperature bay handle (float t)))
    this)
  RandomSampler
  (sample! [this]
    (sample! this walker-count))
  (sample! [this n-or-res]
    (in-context
     ctx
     (let-release [res (if (integer? n-or-res)
                         (ge neanderthal-factory DIM n-or-res {:raw true})
                         n-or-res)]
       ;; the result stays on the device, written straight into the cuda-float matrix (G/:371-389)
       (ok! (.bay_sample bay handle (long (ncols res)) (dev-ptr res) 1))
       res)))
  MCMC
  (init! [this seed]
    (ok! (.bay_init bay handle (int seed)))
    this)
  (init-position! [this position]
    (in-context ctx (ok! (.bay_init_position_from bay handle (.-handle ^B200Stretch position))))
    this)
  (init-position! [this seed limits]
    ;; limits: host 2 x DIM matrix, column-major = (lo_d, hi_d) pairs (G/:409-418)
    (in-context ctx (ok! (.bay_init_position_uniform bay handle (int seed) (host-floats limits))))
    this)
  (burn-in! [this n a]
    (in-context ctx (ok! (.bay_burn_in bay handle (long n) (float a))))
    this)
  (anneal! [this schedule n a]
    ;; the schedule fn i -> T is evaluated on the JVM; the library turns it into 1/T per step (G/:430-440, 366)
    (let [temps (float-array (map #(float (schedule %)) (range n)))]
      (in-context ctx (ok! (.bay_anneal bay handle temps (long n) (float a))))
      this))
  (acc-rate! [this a]
    (in-context
     ctx
     (let [out (DoubleByReference.)]
       (ok! (.bay_acc_rate bay handle (float a) out))
       (.getValue out))))
  (run-sampler! [this n a]
    (in-context
     ctx
     (let [acc (DoubleByReference.) lag (LongByReference.)
           tau (float-array DIM) mean (float-array DIM) sigma (float-array DIM)]
       (ok! (.bay_run_sampler bay handle (long n) (float a) acc tau mean sigma lag))
       {:acceptance-rate (.getValue acc)
        :a a
        :autocorrelation (->Autocorrelation (fv tau) (fv mean) (fv sigma) n (.getValue lag))})))
  EstimateEngine
  (histogram [this]
    (histogram! this 1))
  (histogram! [this cycles]
    (in-context
     ctx
     (let [lim (float-array (* 2 DIM)) pdf (float-array (* WGS DIM)) ranks (float-array (* WGS DIM))]
       (ok! (.bay_histogram bay handle (long cycles) lim pdf ranks))
       (->Histogram (fge 2 DIM lim) (fge WGS DIM pdf) (fge WGS DIM ranks)))))
  Location
  (mean [_]
    (in-context
     ctx
     (let [out (float-array DIM)]
       (ok! (.bay_mean bay handle out))
       (let-release [res (vctr neanderthal-factory DIM)]
         (transfer! out res)))))
  Spread
  (variance [_]
    (in-context
     ctx
     (let [out (float-array DIM)]
       (ok! (.bay_variance bay handle out))
       (let-release [res (vctr neanderthal-factory DIM)]
         (transfer! out res)))))
  (sd [this]
    (in-context
     ctx
     (let [out (float-array DIM)]
       (ok! (.bay_sd bay handle out))
       (let-release [res (vctr neanderthal-factory DIM)]
         (transfer! out res))))))

(defn mix-on-device!
  "mix! (uncomplicate.bayadera.mcmc/mix!, mcmc.clj:66-101) in ONE boundary crossing for the three built-in cooling
  schedules (:minus-n, :sqrt-n, [:pow-n p]); same options and result map.  uncomplicate.bayadera.mcmc/mix! itself keeps
  working unchanged through the protocol methods above (~70 calls)."
  ([^B200Stretch samp {:keys [step dimension-power schedule a min-acc-rate max-acc-rate]
                       :or {step 64 dimension-power 0.8 schedule :minus-n a 2.0 min-acc-rate 0.2 max-acc-rate 0.5}}]
   (let [[kind power] (cond (= schedule :minus-n) [0 1.0] (= schedule :sqrt-n) [1 0.5] :else [2 (double (second schedule))])
         out-a (DoubleByReference.) r1 (DoubleByReference.) r2 (DoubleByReference.)]
     (in-context (.-ctx samp)
                 (ok! (.bay_mix bay (.-handle samp) (long step) (double dimension-power) (int kind) (double power)
                                (double a) (double min-acc-rate) (double max-acc-rate) out-a r1 r2)))
     {:a (.getValue out-a) :acc-rate (.getValue r1) :acc-rate-2.0 (.getValue r2)}))
  ([samp]
   (mix-on-device! samp nil)))

;; ============================ Sampler factory (G/:543-610) ===================

(deftype B200StretchFactory [ctx ^Pointer modl neanderthal-factory model ^long DIM ^long WGS]
  Releaseable
  (release [_]
    (in-context ctx (ok! (.bay_model_release bay modl)))
    true)
  SamplerFactory
  (create-sampler [_ seed walker-count params]
    (in-context
     ctx
     ;; params is a cuda-float vector [data || hyperparams] that the sampler BORROWS, like the reference (G/:558);
     ;; a walker count that is not a multiple of 2*WGS comes back as BAY_EINVAL_WALKERS -> IllegalArgumentException
     ;; with the reference's text (G/:609-610)
     (let [out (PointerByReference.)]
       (ok! (.bay_sampler_create_dev bay modl (int seed) (long walker-count)
                                     (+ (device-pointer (buffer params)) (* Float/BYTES (long (offset params))))
                                     (long (dim params)) out))
       (->B200Stretch ctx (.getValue out) neanderthal-factory model DIM WGS (long walker-count))))))

;; ============================ Engine constructors (G/:640-757) ===============

(defn b200-dataset-engine [ctx ^Pointer eng WGS]
  (->B200DatasetEngine ctx eng (long WGS)))

(defn b200-acor-engine [ctx ^Pointer eng]
  (->B200AcorEngine ctx eng))

(defn b200-distribution-engine [ctx ^Pointer eng model]
  (in-context ctx (->B200DistributionEngine ctx (compile-model eng model (logpdf model)) model)))

(defn b200-likelihood-engine [ctx ^Pointer eng model]
  (in-context
   ctx
   (let [out (PointerByReference.)
         srcs (into-array String (concat (source model) [(likelihood-source model)]))]
     ;; dimension / params-size: a likelihood is evaluated on points of the prior it will be paired with; with
     ;; params-size 0 the whole vector handed to log-density is data (G/:118-127)
     (ok! (.bay_model_compile bay eng srcs (alength srcs) "bay_loglik_engine"
                              (int (or (:dimension model) 1)) (int 0) (int FAST-MATH) out))
     (->B200LikelihoodEngine ctx (.getValue out) model))))

(defn b200-direct-sampler-engine [ctx ^Pointer eng model]
  (if-let [family (direct-families (:name model))]
    (->B200DirectSamplerEngine ctx eng (long family))
    (dragan-says-ex "The B200 engine has direct samplers for uniform, gaussian, exponential and erlang only."
                    {:model (:name model)})))

(defn b200-stretch-factory [ctx ^Pointer eng neanderthal-factory model WGS]
  (in-context
   ctx
   (->B200StretchFactory ctx (compile-model eng model (mcmc-logpdf model)) neanderthal-factory model
                         (long (dimension model)) (long WGS))))

;; =========================== Bayadera factory (G/:761-807) ===================

(defrecord B200BayaderaFactory [ctx hstream ^Pointer eng ^long compute-units ^long WGS
                                neanderthal-factory dataset-eng acor-eng]
  Releaseable
  (release [_]
    (in-context
     ctx
     (release dataset-eng)
     (release acor-eng)
     (release neanderthal-factory)
     (ok! (.bay_engine_release bay eng))
     true))
  na/MemoryContext
  (compatible? [_ o]
    (or (satisfies? DeviceModel o) (na/compatible? neanderthal-factory o)))
  na/FactoryProvider
  (factory [_]
    neanderthal-factory)
  (native-factory [_]
    (na/native-factory neanderthal-factory))
  EngineFactory
  (likelihood-engine [_ model]
    (b200-likelihood-engine ctx eng model))
  (distribution-engine [_ model]
    (b200-distribution-engine ctx eng model))
  (direct-sampler-engine [_ model]
    (b200-direct-sampler-engine ctx eng model))
  (dataset-engine [_]
    dataset-eng)
  (mcmc-factory [_ model]
    (b200-stretch-factory ctx eng neanderthal-factory model WGS))
  (processing-elements [_]
    (* compute-units WGS)))

(defn b200-bayadera-factory
  "Twin of gtx-bayadera-factory (G/:791-807): same arities.  The library adopts the ClojureCUDA context that is
  current inside (in-context ctx ...) — bay_engine_create_current — and enqueues on hstream, so Neanderthal /
  ClojureCUDA work on the same stream stays ordered and cuda-float buffers of ctx can be handed across as they are."
  ([ctx hstream compute-units WGS]
   (in-context
    ctx
    (let [out (PointerByReference.)]
      (ok! (.bay_engine_create_current bay (device-pointer hstream) (int WGS) out))
      (let [eng (.getValue out)]
        (let-release [neanderthal-factory (cuda-float ctx hstream)
                      dataset-eng (b200-dataset-engine ctx eng WGS)
                      acor-eng (b200-acor-engine ctx eng)]
          (->B200BayaderaFactory ctx hstream eng compute-units WGS neanderthal-factory dataset-eng acor-eng))))))
  ([ctx hstream]
   (in-context
    ctx
    (let [dev (ctx-device)]
      (b200-bayadera-factory ctx hstream (multiprocessor-count dev) (max-block-dim-x dev))))))
