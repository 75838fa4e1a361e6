"""Multi-GPU scenario sharding: one process per GPU (torchrun), each rank owns a contiguous block of scenarios -- the
reference's thread fan-out (job_dispatch.hpp:131-134, 142-172) replaced by a static partition with no data-path collective
(SURVEY.md section 8e).  torch.distributed is only plumbing: barrier, max-over-ranks of timings, gathering the small
per-scenario status / iteration arrays so that rank 0 can raise one batch error for the whole job."""
import numpy as np


def scenario_block(n_scenarios: int, rank: int, world_size: int):
    """Contiguous block [begin, end) of rank `rank`: ceil-partition like n_scn / n_gpu, empty blocks allowed."""
    per = -(-n_scenarios // world_size) if world_size > 0 else n_scenarios
    begin = min(rank * per, n_scenarios)
    return begin, min(begin + per, n_scenarios)


def slice_update(update_data: dict, begin: int, end: int):
    """Scenario slice of a batch update dataset (dense 2-D arrays or {"data", "indptr"} sparse buffers)."""
    out = {}
    for comp, val in update_data.items():
        if isinstance(val, dict):
            indptr = np.asarray(val["indptr"])
            lo, hi = int(indptr[begin]), int(indptr[end])
            out[comp] = {"data": val["data"][lo:hi], "indptr": indptr[begin:end + 1] - lo}
        else:
            out[comp] = val[begin:end]
    return out


def gather_scenario_array(local: np.ndarray, n_scenarios: int, dist=None):
    """All ranks contribute their block of a per-scenario array; every rank gets the full array (status, n_iter)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch

    world = dist.get_world_size()
    per = -(-n_scenarios // world)
    device = "cuda" if dist.get_backend() == "nccl" else "cpu"
    padded = torch.zeros(per, dtype=torch.int64, device=device)
    padded[: len(local)] = torch.as_tensor(np.asarray(local, dtype=np.int64), device=device)
    parts = [torch.zeros(per, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat(parts)[:n_scenarios].cpu().numpy().astype(local.dtype)


def calculate_power_flow_sharded(model, update_data: dict, dist=None, calculate=None, **kwargs):
    """Run `model.calculate_power_flow` on this rank's block of scenarios.

    Returns (block, local_result, status_all, n_iter_all).  Output arrays stay local to the rank (they are large; the
    caller writes them to its own slice of the job's output), only status / iteration counts are gathered."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    n_scenarios = next(len(v["indptr"]) - 1 if isinstance(v, dict) else len(v) for v in update_data.values())
    begin, end = scenario_block(n_scenarios, rank, world)
    local_update = slice_update(update_data, begin, end)
    calculate = calculate or (lambda upd, **kw: _run(model, upd, **kw))
    result, status, n_iter = calculate(local_update, **kwargs) if end > begin else ({}, np.zeros(0, np.int32), np.zeros(0, np.int32))
    status_all = gather_scenario_array(status, n_scenarios, dist)
    n_iter_all = gather_scenario_array(n_iter, n_scenarios, dist)
    return (begin, end), result, status_all, n_iter_all


def _run(model, update, **kwargs):
    kwargs.setdefault("continue_on_batch_error", True)
    result = model.calculate_power_flow(update_data=update, **kwargs)
    return result, model.status.copy(), model.n_iter.copy()


def bind_process_to_device_cpus(device: int):
    """One process per GPU: keep this process (and therefore the page-locked host buffers it allocates, first touch) on the CPUs
    that are local to its GPU (NVML's affinity mask = the NUMA node the GPU's PCIe root hangs off).  With 8 ranks each delivering
    hundreds of MB of output structs per step, buffers on the far socket put every transfer across the inter-socket link.
    Returns the CPU list it bound to, or None when NVML gives no usable mask (single-socket hosts, VMs) -- then nothing changes."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[device]) if visible and all(v.strip().isdigit() for v in visible.split(",")) else device
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:  # noqa: BLE001 -- binding is an optimisation, never a requirement
        return None
