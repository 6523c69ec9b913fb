import ctypes as C, os, sys, torch, torch.distributed as dist
from pathlib import Path
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = C.CDLL(str(Path(__file__).resolve().parent / "libipc_probe.so"))
buf = C.create_string_buffer(64)
n = L.probe_init(local, buf)
assert n == 64, n
mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
allh = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allh, mine)
blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in allh)
rc = L.probe_open(rank, world, blob)
assert rc == 0
dist.barrier()
sums = (C.c_double * 2)(); ms = C.c_double()
rc = L.probe_run(rank, world, 1000, sums, C.byref(ms))
exp0 = sum(1000.0 * r + 5 for r in range(world)); exp1 = sum((r + 1) * 1000 for r in range(world))
print(f"rank {rank}: status {rc} peer-read sum {sums[0]} (want {exp0}) allreduce {sums[1]} (want {exp1}) {ms.value*1e3:.2f} us per device allreduce+barrier", flush=True)
dist.barrier()
dist.destroy_process_group()
