"""usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer/pingpong.py
Builds tools/peer/pingpong.cu on the fly is NOT possible on the GPU box path-wise? (nvcc exists there too); we prebuild here."""
import ctypes as C, os, sys, torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
here = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(here, "libpingpong.so"))
lib.pingpong.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]
try:
    import torch.distributed._symmetric_memory as symm
    buf = symm.empty(64, dtype=torch.int64, device=torch.device("cuda", local))
    hdl = symm.rendezvous(buf, group=dist.group.WORLD)
    ptrs = list(hdl.buffer_ptrs)
    print(f"[rank {rank}] symmetric memory ok: ptrs {[hex(p) for p in ptrs]}", flush=True)
except Exception as e:
    print(f"[rank {rank}] symmetric memory unavailable: {type(e).__name__}: {e}", flush=True)
    dist.destroy_process_group(); sys.exit(0)
buf.zero_(); torch.cuda.synchronize(); dist.barrier()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
iters = 20000
peer = ptrs[1 - rank]
for rep in range(2):
    buf.zero_(); torch.cuda.synchronize(); dist.barrier()
    rc = lib.pingpong(C.c_void_p(ptrs[rank]), C.c_void_p(peer), rank, iters, 4_000_000_000, C.c_void_p(out.data_ptr()),
                      C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    cyc, ok = int(out[0]), int(out[1])
    print(f"[rank {rank}] rep {rep}: rc={rc} ok={ok} cycles/round-trip {cyc / iters:.0f}  (~{cyc / iters / 1.9e3:.2f} us at 1.9 GHz)", flush=True)
dist.barrier(); dist.destroy_process_group()
