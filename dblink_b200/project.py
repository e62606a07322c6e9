"""Project: a dblink configuration bound to its data (Project.scala:32-230, ProjectSteps.scala:53-84).

Same HOCON surface as the reference (`dblink.data.*`, `dblink.partitioner`, `dblink.steps[*]`, ...); the sample
step runs on the GPU engine, summarize / evaluate run on the host from linkage-chain.parquet.
"""
import os
import shutil

import numpy as np

from . import analysis, analysis_arrays, config as hocon, sampler as chain, state_io, writers
from .engine import GibbsEngine, KDTreePartitioner
from .records import (Attribute, RecordsCache, SimilarityFn, build_cache_from_columns, read_csv,  # noqa: F401
                      read_csv_columns)

SUPPORTED_METRICS = ("pairwise", "cluster")  # ProjectStep.scala:36
SUPPORTED_QUANTITIES = ("cluster-size-distribution", "partition-sizes", "shared-most-probable-clusters")  # :37


class Project:
    def __init__(self, cfg, base_dir="."):
        g = cfg.get
        self.cfg = cfg
        self.data_path = os.path.join(base_dir, cfg.get_string("dblink.data.path"))
        self.output_path = os.path.join(base_dir, cfg.get_string("dblink.outputPath"))
        self.rec_id_attribute = cfg.get_string("dblink.data.recordIdentifier")
        self.file_id_attribute = g("dblink.data.fileIdentifier", None)
        self.ent_id_attribute = g("dblink.data.entityIdentifier", None)
        self.null_value = cfg.get_string("dblink.data.nullValue")
        self.random_seed = cfg.get_int("dblink.randomSeed")
        self.population_size = g("dblink.populationSize", None)
        self.expected_max_cluster_size = int(g("dblink.expectedMaxClusterSize", 10))
        self.matching_attributes = []
        for c in cfg.get_list("dblink.data.matchingAttributes"):  # Project.parseMatchingAttributes, :201-216
            sf = c["similarityFunction"]
            if sf["name"] == "ConstantSimilarityFn":
                fn = SimilarityFn("ConstantSimilarityFn")
            elif sf["name"] == "LevenshteinSimilarityFn":
                fn = SimilarityFn("LevenshteinSimilarityFn", float(sf["parameters"]["threshold"]),
                                  float(sf["parameters"]["maxSimilarity"]))
            else:
                raise hocon.ConfigError("similarityFunction.name: unsupported value")
            self.matching_attributes.append(Attribute(c["name"], fn, float(c["distortionPrior"]["alpha"]),
                                                      float(c["distortionPrior"]["beta"])))
        p = cfg.get_config("dblink.partitioner")  # Project.parsePartitioner, :218-229
        if p.get_string("name") != "KDTreePartitioner":
            raise hocon.ConfigError("partitioner.name: unsupported value")
        names = [a.name for a in self.matching_attributes]
        self.num_levels = p.get_int("parameters.numLevels")
        self.partition_attribute_ids = [names.index(n) for n in p.get_list("parameters.matchingAttributes")]
        self._loaded = None

    @classmethod
    def from_file(cls, path):
        base = os.getcwd()
        return cls(hocon.parse_file(path), base)

    # ---- description (what Run.main writes to run.txt, Run.scala:38-43) ----------------------------
    def mk_string(self):
        """Project.mkString (Project.scala:58-96)."""
        def sim(fn):
            if fn.name == "ConstantSimilarityFn":
                return "ConstantSimilarityFn"
            return f"LevenshteinSimilarityFn(threshold={fn.threshold}, maxSimilarity={fn.max_similarity})"

        L = ["Data settings", "-------------", f"  * Using data files located at '{self.data_path}'",
             f"  * The record identifier attribute is '{self.rec_id_attribute}'",
             f"  * The file identifier attribute is '{self.file_id_attribute}'" if self.file_id_attribute
             else "  * There is no file identifier",
             f"  * The entity identifier attribute is '{self.ent_id_attribute}'" if self.ent_id_attribute
             else "  * There is no entity identifier",
             "  * The matching attributes are " + ", ".join(f"'{a.name}'" for a in self.matching_attributes), "",
             "Hyperparameter settings", "-----------------------"]
        for i, a in enumerate(self.matching_attributes):
            L.append(f"  * '{a.name}' (id={i}) with {sim(a.similarity_fn)} and "
                     f"BetaShapeParameters(alpha={a.alpha}, beta={a.beta})")
        L += [f"  * Size of latent population is {self.population_size}", "",
              "Partition function settings", "---------------------------"]
        if self.num_levels == 0:
            L.append("  * KDTreePartitioner(numLevels=0)")
        else:
            L.append(f"  * KDTreePartitioner(numLevels={self.num_levels}, attributeIds="
                     f"[{','.join(str(i) for i in self.partition_attribute_ids)}])")
        L += ["", "Project settings", "----------------", f"  * Using randomSeed={self.random_seed}",
              f"  * Using expectedMaxClusterSize={self.expected_max_cluster_size}",
              f"  * Saving Markov chain and complete final state to '{self.output_path}'",
              "  * Sweeps run on the CUDA device of this process (libdblink_b200); there are no Spark checkpoints"]
        return "\n".join(L) + "\n"

    def steps_mk_string(self):
        """ProjectSteps.mkString (ProjectSteps.scala:38-45) with the step descriptions of ProjectStep.scala."""
        L = ["Scheduled steps", "---------------"]
        braces = lambda xs: "{" + ", ".join(f"'{x}'" for x in xs) + "}"
        for name, prm in self.steps():
            if name == "sample":
                src = "saved state" if prm["resume"] else "new initial state"
                L.append(f"  * SampleStep: Evolving the chain from {src} with sampleSize={prm['sample_size']}, "
                         f"burninInterval={prm['burnin_interval']}, thinningInterval={prm['thinning_interval']} and "
                         f"sampler={prm['sampler']}")
            elif name == "summarize":
                L.append(f"  * SummarizeStep: Calculating summary quantities {braces(prm['quantities'])} along the "
                         f"chain for iterations >= {prm['lower_iteration_cutoff']}")
            elif name == "evaluate":
                if prm["use_existing_smpc"]:
                    L.append(f"  * EvaluateStep: Evaluating saved sMPC clusters using {braces(prm['metrics'])} metrics")
                else:
                    L.append(f"  * EvaluateStep: Evaluating sMPC clusters (computed from the chain for iterations >= "
                             f"{prm['lower_iteration_cutoff']}) using {braces(prm['metrics'])} metrics")
            else:
                L.append("  * CopyFilesStep: Copying {" + ", ".join(prm["file_names"]) + "} to destination "
                         + prm["destination_path"])
        return "\n".join(L)

    def write_run_txt(self):
        """run.txt under outputPath: the project and its scheduled steps (Run.scala:38-43)."""
        os.makedirs(self.output_path, exist_ok=True)
        with open(os.path.join(self.output_path, "run.txt"), "w") as fh:
            fh.write(self.mk_string())
            fh.write("\n" + self.steps_mk_string())

    # ---- data ------------------------------------------------------------------------------------
    def load(self):
        if self._loaded is None:
            names = [a.name for a in self.matching_attributes]
            # columnar path (pyarrow): same cache / value ids as read_csv + RecordsCache.build + transform_records
            rec_ids, files, columns, ent_ids = read_csv_columns(self.data_path, self.rec_id_attribute, names,
                                                                self.file_id_attribute, self.ent_id_attribute,
                                                                self.null_value)
            cache, x, f = build_cache_from_columns(columns, files, self.matching_attributes,
                                                   self.expected_max_cluster_size)
            self._loaded = {"rec_ids": rec_ids.to_pylist(), "ent_ids": None if ent_ids is None else ent_ids.to_pylist(),
                            "cache": cache, "x": x, "file": f}
        return self._loaded

    def true_clusters(self):
        d = self.load()
        if d["ent_ids"] is None:
            return None
        return analysis.membership_to_clusters(d["rec_ids"], d["ent_ids"])

    def true_labels(self):
        """Ground truth as a function: record ids (pyarrow array, the chain's dictionary) -> int labels in that
        order; None when the data has no entity id column."""
        d = self.load()
        if d["ent_ids"] is None:
            return None

        def lookup(record_ids):
            import pyarrow as pa
            import pyarrow.compute as pc

            pos = pc.index_in(record_ids, value_set=pa.array([str(r) for r in d["rec_ids"]], pa.string()))
            if pos.null_count:
                raise ValueError("the chain mentions record ids that are not in the data")
            _, lab = np.unique(np.asarray([str(e) for e in d["ent_ids"]], dtype=object), return_inverse=True)
            return lab[pos.to_numpy(zero_copy_only=False)]

        return lookup

    def _new_engine(self):
        d = self.load()
        cache = d["cache"]
        return GibbsEngine(cache.indexes, [a.alpha for a in self.matching_attributes],
                           [a.beta for a in self.matching_attributes], None, self.random_seed, len(cache.file_ids))

    def fingerprint(self):
        d = self.load()
        return state_io.model_fingerprint(d["cache"].indexes, d["x"], d["file"],
                                          [a.alpha for a in self.matching_attributes],
                                          [a.beta for a in self.matching_attributes],
                                          extra=(int(self.random_seed), int(self.population_size or 0),
                                                 int(self.num_levels), tuple(self.partition_attribute_ids)))

    def generate_initial_state(self):
        """Project.generateInitialState (:130-145) -> a GibbsEngine at iteration 0."""
        d = self.load()
        eng = self._new_engine()
        eng.init_state(d["x"], d["file"], int(self.population_size or 0))
        part = KDTreePartitioner(self.num_levels, self.partition_attribute_ids).fit(eng.download_state()["y"])
        eng.set_partitioner(part)
        eng._partitioner_keepalive = part
        return eng

    def saved_state(self):
        """Project.savedState (:113-128): the engine restored from `state.npz` under outputPath, or None.  The
        partition function is re-fitted on the deterministic initial entity values, exactly as the original run
        fitted it, so the resumed chain continues as if it had never stopped; the fingerprint covers the data, the
        tables, the priors, randomSeed, populationSize and the partitioner settings, so a state saved under a
        different configuration is refused instead of being continued with another key / partition function."""
        if not state_io.saved_state_exists(self.output_path):
            return None
        st = state_io.load_state(self.output_path, self.fingerprint())
        d = self.load()
        eng = self.generate_initial_state()
        eng.upload_state(d["x"], d["file"], st["z"], st["link"], st["y"], st["theta"], st["iteration"])
        return eng

    def save_state(self, eng):
        state_io.save_state(eng, self.output_path, self.fingerprint(), self.random_seed)

    # ---- steps (ProjectSteps.parseSteps) -------------------------------------------------------------
    def steps(self):
        out = []
        for st in self.cfg.get_list("dblink.steps"):
            prm = st.get("parameters", {})
            name = st["name"]
            if name == "sample":
                out.append(("sample", dict(sample_size=int(prm["sampleSize"]),
                                           burnin_interval=int(prm.get("burninInterval", 0)),
                                           thinning_interval=int(prm.get("thinningInterval", 1)),
                                           resume=bool(prm.get("resume", True)), sampler=prm.get("sampler", "PCG-I"))))
            elif name == "summarize":
                q = list(prm["quantities"])
                if not q or any(x not in SUPPORTED_QUANTITIES for x in q):
                    raise ValueError(f"quantities must be one of {SUPPORTED_QUANTITIES}.")
                out.append(("summarize", dict(lower_iteration_cutoff=int(prm.get("lowerIterationCutoff", 0)),
                                              quantities=q)))
            elif name == "evaluate":
                m = list(prm["metrics"])
                if not m or any(x not in SUPPORTED_METRICS for x in m):
                    raise ValueError(f"metrics must be one of {SUPPORTED_METRICS}.")
                out.append(("evaluate", dict(lower_iteration_cutoff=int(prm.get("lowerIterationCutoff", 0)), metrics=m,
                                             use_existing_smpc=bool(prm.get("useExistingSMPC", False)))))
            elif name == "copy-files":
                out.append(("copy-files", dict(file_names=list(prm["fileNames"]),
                                               destination_path=prm["destinationPath"],
                                               overwrite=bool(prm.get("overwrite", False)),
                                               delete_source=bool(prm.get("deleteSource", False)))))
            else:
                raise hocon.ConfigError("steps.name: unsupported step")
        return out

    def execute(self, log=print):
        eng = None
        results = {}
        for name, prm in self.steps():
            if name == "sample":
                if prm["resume"]:  # ProjectStep.scala:47-52
                    eng = eng or self.saved_state() or self.generate_initial_state()
                else:
                    eng = self.generate_initial_state()
                d = self.load()
                log(f"SampleStep: sampleSize={prm['sample_size']} burninInterval={prm['burnin_interval']} "
                    f"thinningInterval={prm['thinning_interval']} sampler={prm['sampler']}")
                chain.sample(eng, d["rec_ids"], [a.name for a in self.matching_attributes], prm["sample_size"],
                             self.output_path, prm["burnin_interval"], prm["thinning_interval"], sampler=prm["sampler"])
                self.save_state(eng)  # Sampler.scala:120
            elif name == "summarize":
                # array implementations (analysis_arrays): same quantities as analysis.py, no loop over clusters
                ch = analysis_arrays.read_chain_arrays(os.path.join(self.output_path, "linkage-chain.parquet"),
                                                       prm["lower_iteration_cutoff"])
                for q in prm["quantities"]:
                    if q == "cluster-size-distribution":
                        writers.save_cluster_size_distribution(analysis_arrays.cluster_size_distribution(ch),
                                                               self.output_path)
                    elif q == "partition-sizes":
                        writers.save_partition_sizes(analysis_arrays.partition_sizes(ch), self.output_path)
                    else:
                        labels = analysis_arrays.shared_most_probable_clusters(ch)
                        self._save_smpc(analysis_arrays.labels_to_clusters(labels, ch.record_ids))
            elif name == "evaluate":
                true_labels = self.true_labels()
                if true_labels is None:
                    raise ValueError("Ground truth entity ids are required for evaluation")  # ProjectStep.scala:65
                ch = analysis_arrays.read_chain_arrays(os.path.join(self.output_path, "linkage-chain.parquet"),
                                                       prm["lower_iteration_cutoff"])
                smpc_path = os.path.join(self.output_path, "shared-most-probable-clusters.csv")
                if prm["use_existing_smpc"] and os.path.exists(smpc_path):  # ProjectStep.scala EvaluateStep
                    labels = self._read_smpc_labels(smpc_path, ch.record_ids)
                else:
                    labels = analysis_arrays.shared_most_probable_clusters(ch)
                    self._save_smpc(analysis_arrays.labels_to_clusters(labels, ch.record_ids))
                truth = true_labels(ch.record_ids)
                text = []
                for m in prm["metrics"]:
                    if m == "pairwise":
                        results["pairwise"] = analysis_arrays.pairwise_metrics(labels, truth)
                        text.append(analysis.format_pairwise(results["pairwise"]))
                    else:
                        results["cluster"] = analysis_arrays.adjusted_rand_index(labels, truth)
                        text.append(analysis.format_cluster(results["cluster"]))
                with open(os.path.join(self.output_path, "evaluation-results.txt"), "w") as fh:
                    fh.write("\n".join(text) + "\n")
            elif name == "copy-files":
                os.makedirs(prm["destination_path"], exist_ok=True)
                for fn in prm["file_names"]:
                    src = os.path.join(self.output_path, fn)
                    if os.path.exists(src):
                        dst = os.path.join(prm["destination_path"], os.path.basename(fn))
                        if os.path.exists(dst) and not prm["overwrite"]:
                            continue
                        (shutil.copytree if os.path.isdir(src) else shutil.copy)(src, dst)
                        if prm["delete_source"]:
                            (shutil.rmtree if os.path.isdir(src) else os.remove)(src)
        return results

    @staticmethod
    def _read_smpc_labels(path, record_ids):
        """Labels (aligned with record_ids) from a saved shared-most-probable-clusters.csv: one cluster per line."""
        pos = {r: i for i, r in enumerate(record_ids.to_pylist())}
        labels = np.full(len(pos), -1, np.int64)
        with open(path) as fh:
            for c, line in enumerate(fh):
                for rid in (t.strip() for t in line.split(",")):
                    if rid:
                        labels[pos[rid]] = c
        if (labels < 0).any():
            raise ValueError("shared-most-probable-clusters.csv does not cover every record of the chain")
        return labels

    def _save_smpc(self, clusters):
        with open(os.path.join(self.output_path, "shared-most-probable-clusters.csv"), "w") as fh:
            for c in clusters:
                fh.write(", ".join(sorted(c)) + "\n")
