;;   Entry namespace of the B200 engine: the twin of uncomplicate.bayadera.cuda
;;   (/root/reference/src/clojure/uncomplicate/bayadera/cuda.clj:9-31).  User code changes ONE require:
;;
;;     (require '[uncomplicate.bayadera.b200 :refer [with-default-bayadera]])     ; was uncomplicate.bayadera.cuda
;;
;;   and keeps using uncomplicate.bayadera.core / mcmc / library unchanged.  The model sources are the reference's own
;;   uncomplicate/bayadera/internal/device/cuda/*.cu resource files (compiled by NVRTC for sm_100a inside
;;   libbayadera_b200.so instead of by ClojureCUDA).

(ns uncomplicate.bayadera.b200
  (:require [uncomplicate.clojurecuda.core :refer [with-default current-context default-stream]]
            [uncomplicate.bayadera
             [core :refer [with-bayadera *bayadera-factory*]]
             [library :refer [with-library]]]
            [uncomplicate.bayadera.internal.device
             [models :as models]
             [b200 :refer [b200-bayadera-factory]]]))

(def source-library (models/source-library "uncomplicate/bayadera/internal/device/cuda/%s.cu"))

(def device-library (partial models/device-library source-library))

(defmacro with-default-library [factory & body]
  `(with-library (device-library ~factory)
     ~@body))

(defmacro with-default-bayadera [& body]
  `(with-default
     (with-bayadera b200-bayadera-factory [(current-context) default-stream]
       (with-default-library *bayadera-factory*
         ~@body))))
