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
                                          [a.beta for a in self.matching_attributes])

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
        fitted it, so the resumed chain continues as if it had never stopped."""
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

    def _save_smpc(self, clusters):
        with open(os.path.join(self.output_path, "shared-most-probable-clusters.csv"), "w") as fh:
            for c in clusters:
                fh.write(", ".join(sorted(c)) + "\n")
