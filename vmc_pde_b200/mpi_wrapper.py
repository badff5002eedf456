"""Mirror of vmc_fluids/mpi_wrapper.py on torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests).

Same module-level surface: rank, commSize, comm, globNumSamples, global_sum / global_mean /
global_variance / global_covariance, distribute_sampling, first_sample_id, bcast_unknown_size,
get_communication_time.  The reference stages every reduction through host NumPy + MPI.Allreduce
(mpi_wrapper.py:150-163,229-245); here the reduction is a device-side all_reduce of the same values.
The fused TDVP path (tdvp.py) does not call these per quantity: it packs first moments into one
all-reduce and second moments into another.
"""
import time
import numpy as np
import torch
import torch.distributed as dist

globNumSamples = 0     # mpi_wrapper.py:13
myNumSamples = 0
communicationTime = 0.0


def _initialized():
    return dist.is_available() and dist.is_initialized()


class _Comm:
    """Stands in for MPI.COMM_WORLD (mpi_wrapper.py:9-11)."""

    def Get_rank(self):
        return dist.get_rank() if _initialized() else 0

    def Get_size(self):
        return dist.get_world_size() if _initialized() else 1

    def Barrier(self):
        if _initialized():
            dist.barrier()


comm = _Comm()


def __getattr__(name):
    # rank / commSize are module attributes in the reference; keep them live w.r.t. init_process_group
    if name == "rank":
        return comm.Get_rank()
    if name == "commSize":
        return comm.Get_size()
    raise AttributeError(name)


def allreduce_(t):
    """In-place SUM all-reduce of a tensor across ranks (no-op for a single process)."""
    global communicationTime
    if _initialized() and dist.get_world_size() > 1:
        t0 = time.perf_counter()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        communicationTime += time.perf_counter() - t0
    return t


def broadcast_(t, src=0):
    """In-place broadcast of a tensor from rank `src` (no-op for a single process)."""
    global communicationTime
    if _initialized() and dist.get_world_size() > 1:
        t0 = time.perf_counter()
        dist.broadcast(t, src=src)
        communicationTime += time.perf_counter() - t0
    return t


def distribute_sampling(numSamples, localDevices=None, numChainsPerDevice=1):
    """mpi_wrapper.py:68-110, same arithmetic."""
    global globNumSamples
    size, r = comm.Get_size(), comm.Get_rank()
    samplesPerProcess = numSamples // size
    if r < numSamples % size:
        samplesPerProcess += 1
    if localDevices is None:
        globNumSamples = numSamples
        return samplesPerProcess
    numChainsPerProcess = localDevices * numChainsPerDevice

    def spc(spp):
        return (spp + numChainsPerProcess - 1) // numChainsPerProcess

    a = numSamples % size
    globNumSamples = (a * spc(1 + numSamples // size) + (size - a) * spc(numSamples // size)) * numChainsPerProcess
    return spc(samplesPerProcess)


def first_sample_id():
    """mpi_wrapper.py:113-126."""
    size, r = comm.Get_size(), comm.Get_rank()
    mySamples = globNumSamples // size
    firstSampleId = r * mySamples
    if r < globNumSamples % size:
        firstSampleId += r
    else:
        firstSampleId += globNumSamples % size
    return firstSampleId


def shard_range(numSamples):
    """Contiguous global sample range [first, first + n) owned by this rank."""
    size, r = comm.Get_size(), comm.Get_rank()
    base, rem = numSamples // size, numSamples % size
    n = base + (1 if r < rem else 0)
    first = r * base + min(r, rem)
    return first, n


def global_sum(data):
    """mpi_wrapper.py:129-163: sum over the (device, batch) axes and over ranks."""
    res = data.sum(dim=(0, 1)) if data.ndim >= 2 else data.sum()
    res = res.clone()
    return allreduce_(res)


def global_mean(data):
    """mpi_wrapper.py:166-193."""
    return global_sum(data) / globNumSamples


def global_variance(data):
    """mpi_wrapper.py:196-245: mean of |x - mean|^2."""
    mean = global_mean(data)
    d = data - mean
    res = (d.conj() * d).sum(dim=(0, 1)).clone() if data.ndim >= 2 else (d.conj() * d).sum().clone()
    return allreduce_(res) / globNumSamples


def global_covariance(data):
    """mpi_wrapper.py:248-274 with _cov_helper_without_p (:21-25): (1/N) sum_i x_i x_i^H via the DMMA Gram kernel."""
    from . import _kernels
    n = data.shape[0] * data.shape[1]
    S = _kernels.gram_plain(data.reshape(n, data.shape[-1]))
    return allreduce_(S) / globNumSamples


def bcast_unknown_size(data, root=0):
    """mpi_wrapper.py:277-306."""
    if comm.Get_rank() == root:
        if np.asarray(data).dtype != np.float64:
            raise TypeError("Datatype has to be float64.")
    if not (_initialized() and dist.get_world_size() > 1):
        return np.array(data)
    obj = [np.array(data) if comm.Get_rank() == root else None]
    dist.broadcast_object_list(obj, src=root)
    return obj[0]


def get_communication_time():
    """mpi_wrapper.py:309-313."""
    global communicationTime
    t = communicationTime
    communicationTime = 0.0
    return t
