/*
 * dblink_b200.h -- C ABI of the B200-native Gibbs-sweep engine for cleanzr/dblink's record-linkage model.
 *
 * The reference (Scala/Spark, no FFI of its own) has exactly one seam around the hot path; every entry
 * point below names the reference interface it replaces.  Paths are relative to
 * src/main/scala/com/github/cleanzr/dblink/ of cleanzr/dblink @ dc3dd0d; GU = GibbsUpdates.scala.
 *
 * Conventions (JNI/JCuda/ctypes friendly): opaque handles, plain pointers + sizes, int status
 * (0 = ok, negative = error; dbl_last_error() gives the text), no callbacks, no exceptions across the
 * boundary.  The caller owns every host buffer; the library owns every device buffer.  One context per
 * process/GPU; calls on one context are serialised by the caller.  All compute runs on the CUDA device
 * that is current when dbl_ctx_create() is called -- there is no CPU fallback.
 */
#ifndef DBLINK_B200_H
#define DBLINK_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBL_OK 0
#define DBL_ERR_INVALID (-1)   /* bad argument (reference: require(...) -> IllegalArgumentException)            */
#define DBL_ERR_CUDA (-2)      /* CUDA runtime failure / no usable device                                       */
#define DBL_ERR_ZERO_MASS (-3) /* a categorical had zero/non-finite mass (random/IndexNonUniformDiscreteDist.scala:71-79) */
#define DBL_ERR_STATE (-4)     /* call sequence error (e.g. sweep before state upload)                          */

/* sampler = ProjectStep.scala:35,53-58 */
#define DBL_PCG_I 0            /* collapsedEntityIds=false, collapsedEntityValues=true  (default)               */
#define DBL_PCG_II 1           /* collapsedEntityIds=true,  collapsedEntityValues=true                          */
#define DBL_GIBBS 2            /* both false                                                                    */
#define DBL_GIBBS_SEQ 3        /* "Gibbs-Sequential": same conditionals as DBL_GIBBS (the reference's flag only
                                  disables its inverted index, GU:194-196)                                      */

#define DBL_MAX_ATTRS 32

typedef struct dbl_index dbl_index;   /* AttributeIndex  (AttributeIndex.scala:39-104)                          */
typedef struct dbl_kdtree dbl_kdtree; /* KDTreePartitioner / MutableBST (partitioning/KDTreePartitioner.scala)  */
typedef struct dbl_ctx dbl_ctx;       /* State + broadcast RecordsCache/PartitionFunction (State.scala:56-68)   */

/* ---------------------------------------------------------------------------------------------------
 * Model tables.  Replaces AttributeIndex.apply (AttributeIndex.scala:107-127): value ids in sorted-string
 * order, empirical pmf, sparse exp(similarity) rows (computeSimValueIndex :219-231), normalisations
 * (computeSimNormalizations :234-245), cached base pmfs k=0..kmax (getSimNormDist :197-216,
 * RecordsCache.scala:112-113).  similarity: 0 = ConstantSimilarityFn, 1 = LevenshteinSimilarityFn
 * (SimilarityFn.scala:50-107).  `values` need not be sorted; `weights` are the value counts.
 * ------------------------------------------------------------------------------------------------- */
int dbl_index_build(dbl_index **out, const char *const *values, const double *weights, int32_t num_values,
                    int similarity, double threshold, double max_similarity, int32_t kmax);
/* Same object from pre-computed tables (phi = weight/total as in AttributeIndex.scala:114-115). */
int dbl_index_from_tables(dbl_index **out, int32_t num_values, int similarity, const double *probs,
                          const int32_t *rowptr, const int32_t *col, const double *expsim, int32_t kmax);
void dbl_index_free(dbl_index *);
int32_t dbl_index_num_values(const dbl_index *);                 /* AttributeIndex.numValues                 */
int32_t dbl_index_nnz(const dbl_index *);
/* slots of the per-row perfect-hash tables the link kernel probes (32 = the fast instantiations; 0 = none: constant
 * attribute, or rows too long for a table) */
int32_t dbl_index_hash_slots(const dbl_index *);
int32_t dbl_index_value_id(const dbl_index *, const char *value); /* valueIdxOf; -1 when absent              */
const char *dbl_index_value(const dbl_index *, int32_t value_id);
/* copies of the tables (arrays sized num_values, num_values+1, nnz, nnz); any pointer may be NULL */
int dbl_index_tables(const dbl_index *, double *phi /*probabilityOf*/, double *norm /*simNormalizationOf*/,
                     int32_t *rowptr, int32_t *col, double *expsim /*simValuesOf*/);
double dbl_index_exp_sim(const dbl_index *, int32_t v1, int32_t v2); /* expSimOf; NaN when out of range        */
/* SimilarityFn.getSimilarity (SimilarityFn.scala:65-70, 84-96) */
double dbl_similarity(int similarity, const char *a, const char *b, double threshold, double max_similarity);

/* ---------------------------------------------------------------------------------------------------
 * Partition function.  Replaces KDTreePartitioner.fit / getPartitionId
 * (partitioning/KDTreePartitioner.scala:37-62), MutableBST (MutableBST.scala:51-111) and the
 * DomainSplitters (DomainSplitter.scala:43-110).  y = E x A entity value ids, row-major.
 * ------------------------------------------------------------------------------------------------- */
int dbl_kdtree_fit(dbl_kdtree **out, const int32_t *y, int64_t num_entities, int32_t num_attrs,
                   int32_t num_levels, const int32_t *attr_ids, int32_t num_attr_ids);
int dbl_kdtree_from_arrays(dbl_kdtree **out, int32_t num_nodes, const int32_t *attr, const int32_t *kind,
                           const int32_t *split, const int32_t *set_ptr, const int32_t *set_val,
                           const int32_t *leaf_no);
void dbl_kdtree_free(dbl_kdtree *);
int32_t dbl_kdtree_num_nodes(const dbl_kdtree *);
int32_t dbl_kdtree_num_leaves(const dbl_kdtree *); /* PartitionFunction.numPartitions */
int32_t dbl_kdtree_set_len(const dbl_kdtree *);
int dbl_kdtree_export(const dbl_kdtree *, int32_t *attr, int32_t *kind, int32_t *split, int32_t *set_ptr,
                      int32_t *set_val, int32_t *leaf_no);
int32_t dbl_kdtree_partition_id(const dbl_kdtree *, const int32_t *entity_values); /* getPartitionId */

/* ---------------------------------------------------------------------------------------------------
 * Context = model (RecordsCache + PartitionFunction + Parameters broadcast once, State.scala:216-217,312)
 * ------------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t num_attrs;                 /* A <= DBL_MAX_ATTRS                                                */
  int32_t num_files;                 /* F (RecordsCache.fileSizes keys, sorted)                           */
  const dbl_index *const *indexes;   /* A attribute indexes                                               */
  const double *alpha;               /* A: distortionPrior alpha (BetaShapeParameters, package.scala:164) */
  const double *beta;                /* A                                                                 */
  const dbl_kdtree *tree;            /* NULL = a single block                                             */
  uint64_t seed;                     /* dblink.randomSeed                                                 */
  int32_t rank, world_size;          /* block shard owned by this context: blocks b with owner[b]==rank   */
} dbl_model_desc;

/* The CUDA device of the contexts the calling thread creates next (cudaSetDevice), and how many there are: for
 * hosts without CUDA bindings of their own (one JVM driving several GPUs, one thread per context). */
int dbl_set_device(int32_t device);
int32_t dbl_device_count(void);
int dbl_ctx_create(dbl_ctx **out, const dbl_model_desc *desc);
void dbl_ctx_destroy(dbl_ctx *);
const char *dbl_last_error(const dbl_ctx *); /* never NULL; "" when no error */

/* Install / replace the partition function (PartitionFunction.fit happens on the *initial entity values*,
 * State.scala:309-312, so the usual order is: create the context with tree = NULL, dbl_state_init,
 * dbl_state_download(y), dbl_kdtree_fit, dbl_set_partitioner).  The tree is copied to the device; the
 * caller keeps ownership.  NULL = a single block.  If a state is present it is re-partitioned. */
int dbl_set_partitioner(dbl_ctx *, const dbl_kdtree *tree);
int32_t dbl_num_partitions(const dbl_ctx *);

/* Deterministic initial state, State.deterministic (State.scala:205-334): one entity per record (or
 * population_size entities), values copied from the records, missing values drawn from phi, z = (x>=0 &
 * x!=y), theta = prior mean (DistortionProbs.scala:33-43).  x = R x A value ids (-1 missing,
 * RecordsCache.scala:127-130), file = R file ids in [0,F).  population_size <= 0 means R. */
int dbl_state_init(dbl_ctx *, int64_t num_records, const int32_t *x, const int32_t *file,
                   int64_t population_size);
/* Arbitrary state (resume; State.read, State.scala:160-193).  z = R x A bytes, link = R global entity ids,
 * y = E x A value ids, theta = A x F (host).  x, file, z, link, y may be host OR device pointers (unified
 * addressing decides): a multi-GPU caller can stage slices and all-gather them on the device first.
 * x = file = NULL keeps the records the context already holds (they never change along a chain: the reference
 * broadcasts its RecordsCache once); R and E must then be those of the state being replaced. */
int dbl_state_upload(dbl_ctx *, int64_t num_records, int64_t num_entities, const int32_t *x, const int32_t *file,
                     const uint8_t *z, const int32_t *link, const int32_t *y, const double *theta,
                     int64_t iteration);
/* Full state back to the host (State.save, State.scala:122-150); any pointer may be NULL. */
int dbl_state_download(dbl_ctx *, uint8_t *z, int32_t *link, int32_t *y, double *theta, int32_t *block_of_entity);
int64_t dbl_num_records(const dbl_ctx *);
int64_t dbl_num_entities(const dbl_ctx *);
int64_t dbl_iteration(const dbl_ctx *);

/* n_sweeps applications of the Markov transition operator State.nextState (State.scala:78-99):
 * updateDistProbs (GU:305-320) -> updatePartitions/updatePartition (GU:124-211: link draw per record
 * GU:363-466, entity values GU:731-755, distortions GU:324-359, new partition ids GU:206) ->
 * updateSummaryVariables (GU:219-301).  Everything runs on the device, the A x F Beta draws included; the host waits
 * once, at the end of the call.  Works on sharded contexts as well (after dbl_comm_import, see below). */
int dbl_sweep(dbl_ctx *, int sampler, int32_t n_sweeps);

/* The same transition one k-d-tree block at a time, mirroring the reference's per-partition task
 * GibbsUpdates.updatePartition (GU:156-211), for call-for-call comparison with a CPU chain:
 *   dbl_block_sweep_begin   updateDistProbs (GU:305-320): theta for the new iteration; block membership is frozen
 *   dbl_update_block        link draws, entity values, distortions and new partition ids of ONE block (each block
 *                           exactly once per sweep, in any order); rows of other blocks are not touched
 *   dbl_block_sweep_end     the shuffle by new partition id (GU:144) + updateSummaryVariables (GU:219-301)
 * begin + every block + end leaves exactly the state dbl_sweep(ctx, sampler, 1) leaves.  Unsharded contexts only. */
int dbl_block_sweep_begin(dbl_ctx *, int sampler);
int dbl_update_block(dbl_ctx *, int32_t block_id);
int dbl_block_sweep_end(dbl_ctx *);

/* Linkage structure for linkage-chain.parquet (State.getLinkageStructure, State.scala:102-112): record ->
 * entity links and each entity's current partition id. */
int dbl_links_download(dbl_ctx *, int32_t *link_out /*R*/, int32_t *block_of_entity_out /*E*/);

/* SummaryVars (package.scala:116-119) + what DiagnosticsWriter prints (DiagnosticsWriter.scala:39-72). */
typedef struct {
  int64_t iteration;
  int64_t num_isolates;
  double log_likelihood;
  int64_t pairs_scored; /* (record, candidate) pairs visited by link kernels since ctx creation: the block's entity
                           count per record in the dense kernels, the length of the posting list walked in the pruned one */
} dbl_summary_head;
int dbl_summary(dbl_ctx *, dbl_summary_head *head, int64_t *agg_dist /*A*F*/, int64_t *rec_dist /*A+1*/,
                double *theta /*A*F*/);

/* n sweeps enqueued on the context's stream without waiting for them (dbl_sweep = dbl_sweep_async + dbl_sync).
 * Nothing in a sweep needs the host: theta is drawn on the device, sizes are read from device memory, the exchange
 * of a sharded context is peer to peer.  dbl_sync waits, refreshes what dbl_summary / dbl_iteration report and
 * returns the status of the batch.  A sweep that meets a categorical without mass is ABANDONED: the state, theta
 * and the iteration are those before that sweep, later sweeps of the batch are skipped, DBL_ERR_ZERO_MASS is
 * returned (the reference fails the task and no new state exists, IndexNonUniformDiscreteDist.scala:78-79). */
int dbl_sweep_async(dbl_ctx *, int sampler, int32_t n_sweeps);
int dbl_sync(dbl_ctx *);

/* Order-independent fingerprint of the rows this context owns: hash_out[0] over (entity id, values), hash_out[1]
 * over (record id, link, distortion bits).  Sums over ranks (mod 2^64) identify the global state for any number of
 * ranks and any block placement. */
int dbl_state_hash(dbl_ctx *, uint64_t *hash_out /*2*/);

/* ---------------------------------------------------------------------------------------------------
 * Multi-GPU: one context per rank/GPU (dbl_model_desc.rank / world_size), blocks sharded over ranks.  Replaces
 * the one-task-per-partition fan-out (GU:137), the shuffle `.partitionBy(partitioner)` (GU:144), the broadcast of
 * theta (State.scala:84) and the summary accumulators (SummaryAccumulators.scala:54-63).  Every rank initialises
 * the same replicated state (dbl_state_init / dbl_state_upload, dbl_set_partitioner), then dbl_set_block_owners
 * carves out its shard.  Draws are keyed by global ids, so the chain is identical for any number of ranks.
 *
 * Data plane (a) -- peer to peer, inside the library, no host in the sweep:
 *   dbl_comm_export   allocates this rank's communication buffer and describes it in a DBL_COMM_BLOB_BYTES blob
 *                     (CUDA IPC handle); the host layer all-gathers the blobs with whatever it has (sockets, MPI,
 *                     torch.distributed, Spark's driver) -- this is the only thing it ever moves
 *   dbl_comm_import   maps every peer's buffer over NVLink / NVSwitch
 *   dbl_sweep / dbl_sweep_async then run the whole transition on the device: clusters whose new block belongs to
 *   another rank are written straight into that rank's receive buffer by the kernel that finds them, the partial
 *   summaries go into a slot of every peer, one flag barrier per sweep, theta is drawn redundantly on every rank.
 *   Block -> rank placement is re-evaluated on the device every `period` sweeps (dbl_set_rebalance) from the
 *   global block sizes with the LPT rule of partitioning/LPTScheduler.scala:57-76; re-placed blocks migrate as
 *   ordinary cluster messages.
 * Data plane (b) -- host-mediated (no peer access, several nodes): per sweep
 *   dbl_sweep_begin        theta | global summary, links, entity values, distortions of the owned shard; counts of
 *                          entity / record messages for every destination rank
 *   dbl_exchange_pack      messages into caller-provided DEVICE buffers, concatenated by destination rank:
 *                          entity message = 1 + A int32 words [e, y_0..y_{A-1}], record message = 3 words
 *                          [r, e, z bit mask]; the host moves them with one all-to-all (NCCL) each
 *   dbl_exchange_unpack    apply what was received
 *   dbl_partial_summary -> all-reduce on the host (with an error flag) -> dbl_sweep_end(global summary)
 * ------------------------------------------------------------------------------------------------- */
int dbl_set_block_owners(dbl_ctx *, const int32_t *owner_of_block /* numPartitions entries in [0, world); NULL =
                                                                      the table already on the device */);
int dbl_block_owners(dbl_ctx *, int32_t *owner_of_block_out);  /* current table (the device-side LPT may change it) */
#define DBL_COMM_BLOB_BYTES 192
int dbl_comm_export(dbl_ctx *, void *blob_out /* DBL_COMM_BLOB_BYTES */);
int dbl_comm_import(dbl_ctx *, const void *blobs /* world * DBL_COMM_BLOB_BYTES, in rank order */, int32_t world);
/* period = 0 switches the re-placement off; threshold = current makespan / LPT makespan above which the LPT table is
 * adopted (default: every 16 sweeps, 1.03) */
int dbl_set_rebalance(dbl_ctx *, int32_t period, double threshold);
/* cluster messages this rank sent in the last sweep collected by dbl_sync, placements adopted so far */
int dbl_last_exchange(dbl_ctx *, int64_t *ent_msgs, int64_t *rec_msgs, int64_t *replacements);
int dbl_sweep_begin(dbl_ctx *, int sampler, int64_t *ent_msgs_per_dest /*world*/, int64_t *rec_msgs_per_dest /*world*/);
int dbl_exchange_pack(dbl_ctx *, void *ent_buf_dev, void *rec_buf_dev);
int dbl_exchange_unpack(dbl_ctx *, const void *ent_buf_dev, int64_t n_ent_msgs, const void *rec_buf_dev,
                        int64_t n_rec_msgs);
int32_t dbl_summary_words(const dbl_ctx *); /* A*F + (A+1) + 2 */
int dbl_partial_summary(dbl_ctx *, int64_t *counts /*dbl_summary_words*/, double *loglik_without_prior);
int dbl_sweep_end(dbl_ctx *, const int64_t *global_counts, double global_loglik, int32_t failed_somewhere);
/* the rows this rank owns (zeros elsewhere) into caller-provided DEVICE buffers: y int32[E*A], block int32[E],
 * link int32[R], z uint8[R*A]; summing them over ranks (all-reduce) gives the full state on every rank */
int dbl_export_owned_dev(dbl_ctx *, void *y_dev, void *block_dev, void *link_dev, void *z_dev);
/* only the rows this rank owns, compacted, to the HOST: ids + rows (buffers sized for E / R rows); together the
 * ranks' rows are the state (State.save of a distributed state, State.scala:122-150) */
int dbl_download_owned(dbl_ctx *, int64_t *n_ent, int32_t *ent_ids, int32_t *y, int32_t *block_of_entity,
                       int64_t *n_rec, int32_t *rec_ids, int32_t *link, uint8_t *z);
/* which entities / records this rank currently owns (host byte arrays of E and R entries) */
int dbl_owned_masks(dbl_ctx *, uint8_t *ent_owned, uint8_t *rec_owned);

/* The protocol functions of the theta draw, exposed so that they can be checked without a GPU:
 * log / exp built from individually rounded binary64 operations, and updateDistProbs (GU:305-320) itself. */
double dbl_det_log(double x);
double dbl_det_exp(double x);
int dbl_draw_theta(int32_t num_attrs, int32_t num_files, const double *alpha, const double *beta, uint64_t seed,
                   const int64_t *agg_dist /*A*F*/, const int64_t *file_sizes /*F*/, int64_t iteration,
                   double *theta_out /*A*F*/);

/* count of kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t dbl_kernel_launches(const dbl_ctx *);
/* Link-kernel selection: 0 = automatic (PCG-II: TMA-staged dense kernel; PCG-I/Gibbs: scoring pruned through a
 * per-sweep inverted index), 1 = always the generic fallback kernel, 2 = dense TMA kernels for every sampler.
 * All produce identical draws; this exists so tests can cover every kernel. */
int dbl_set_link_mode(dbl_ctx *, int mode);
/* which link kernel a sweep with this sampler launches: 0 generic fallback, 1 dense must-match kernel, 2 index-pruned
 * kernel, 3 k_link_pcg2 (+4: byte-packed constants, +8: 32-slot tables known at compile time).  A model that falls
 * back to the generic kernel runs an order of magnitude slower: benches and tests assert what they expect. */
int dbl_link_kernel(const dbl_ctx *, int sampler);
/* How dbl_sweep / dbl_sweep_async enqueue several sweeps: 0 = automatic (problems small enough to be bound by kernel
 * launches, and the index-pruned samplers at every size, replay a captured CUDA graph of one sweep), 1 = every sweep
 * enqueued kernel by kernel (and timed phase by phase, dbl_link_kernel_ms / dbl_phase_ms), 2 = graphs whenever
 * possible.  Same chain either way. */
int dbl_set_graph_mode(dbl_ctx *, int mode);
/* CUDA-event time (ms) of the last dbl_sweep call, first operation to last operation on the context's stream */
double dbl_last_sweep_ms(const dbl_ctx *);
/* CUDA-event time (ms) spent in the link-scoring kernel since the last call; resets the accumulator */
double dbl_link_kernel_ms(dbl_ctx *, int64_t *launches);
/* CUDA-event time (ms) of the eagerly enqueued sweeps since the last call, by phase: out4 = {link update GU:186-199,
 * entity values + distortions + partial summary GU:201-210, exchange of moved clusters + summary reduction GU:144 (0
 * on one rank), re-layout by block}; returns the number of sweeps summed; resets the accumulators */
int64_t dbl_phase_ms(dbl_ctx *, double *out4);
const char *dbl_version(void);

#ifdef __cplusplus
}
#endif
#endif
