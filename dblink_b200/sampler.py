"""Chain driver: Sampler.sample (Sampler.scala:51-124) over a GibbsEngine / ShardedGibbs."""
import os

from .writers import DiagnosticsWriter, LinkageChainWriter, linkage_structure_arrow

SUPPORTED_SAMPLERS = ("PCG-I", "PCG-II", "Gibbs", "Gibbs-Sequential")  # ProjectStep.scala:35


def sample(engine, record_ids, attribute_names, sample_size, output_path, burnin_interval=0, thinning_interval=1,
           write_buffer_size=10, sampler="PCG-I", population_size=None, on_sample=None):
    """Generates `sample_size` posterior samples by successively applying the transition operator; writes
    linkage-chain.parquet and diagnostics.csv under output_path.  Returns the number of sweeps performed.
    on_sample(summary, parts) sees every recorded sample; parts = {partition id: pyarrow ListArray of clusters}
    (`.to_pylist()` gives the lists of record ids)."""
    if sample_size <= 0:
        raise ValueError("`sampleSize` must be positive.")            # Sampler.scala:61
    if burnin_interval < 0:
        raise ValueError("`burninInterval` must be non-negative.")    # :62
    if thinning_interval <= 0:
        raise ValueError("`thinningInterval` must be positive.")      # :63
    if write_buffer_size <= 0:
        raise ValueError("`writeBufferSize` must be positive.")       # :65
    if sampler not in SUPPORTED_SAMPLERS:
        raise ValueError(f"sampler must be one of {', '.join(SUPPORTED_SAMPLERS)}.")  # ProjectStep.scala:44
    os.makedirs(output_path, exist_ok=True)
    initial_iteration = engine.iteration
    continue_chain = initial_iteration != 0
    pop = population_size if population_size is not None else engine.num_entities
    lw = LinkageChainWriter(os.path.join(output_path, "linkage-chain.parquet"), write_buffer_size, continue_chain)
    dw = DiagnosticsWriter(os.path.join(output_path, "diagnostics.csv"), attribute_names, continue_chain)

    import pyarrow as pa

    ids = pa.array([str(r) for r in record_ids], pa.string())

    def record():
        link, blk = engine.links()
        parts = linkage_structure_arrow(link, blk, ids)  # {partition id: ListArray of clusters}
        s = engine.summary()
        lw.append(s["iteration"], parts)
        dw.write_row(s, pop)
        if on_sample:
            on_sample(s, parts)

    if not continue_chain and burnin_interval == 0:
        record()  # the initial state is a sample (Sampler.scala:84-89)
    count, done = 0, 0
    while count < sample_size:
        engine.sweep(sampler, 1)
        done += 1
        completed = engine.iteration - initial_iteration
        if completed >= burnin_interval and (completed - burnin_interval) % thinning_interval == 0:
            record()
            count += 1
    lw.close()
    dw.close()
    return done
