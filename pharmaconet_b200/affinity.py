"""CPU / NUMA placement of a one-process-per-GPU rank.

The streamed screening path copies the whole library from pinned host memory every pass; with several ranks on one
box the copy of a rank whose pages sit on the other socket crosses the inter-socket link and becomes the slowest
rank of the step. Binding the process to the CPUs NVML reports as local to its GPU *before* the library is pinned
puts the pages on the GPU's NUMA node (first touch). Plumbing only: no effect on results, silently skipped when NVML
is unavailable."""

from __future__ import annotations

import os


def bind_to_gpu(gpu_index: int, world: int = 1) -> list[int] | None:
    """Restrict this process to the CPUs local to `gpu_index`. Returns the CPU list, or None if nothing was done.
    When NVML reports no locality (a VM without NUMA information: every CPU is "local" to every GPU) and `world` > 1,
    the allowed CPUs are cut into `world` equal slices and rank `gpu_index` takes its own, so that the ranks' copy /
    launch threads at least do not share cores."""
    try:
        import pynvml

        pynvml.nvmlInit()
        try:
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus or len(cpus) == len(allowed):
            if world <= 1 or len(allowed) < world:
                return None
            ordered = sorted(allowed)
            per = len(ordered) // world
            cpus = ordered[(gpu_index % world) * per : (gpu_index % world + 1) * per]
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001 - placement is an optimisation, never a failure
        return None
