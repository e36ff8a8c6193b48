"""Developer probe (2+ GPUs, torchrun): where does the time of tfft_mg_exec go?  Times, per rank and concurrently on
all ranks: the three phases of the six-step separately (tfft_mg_exec_phase, host barriers in between), a copy-engine
peer copy and a plain SM copy kernel storing into the peer (tfft_copy_runs) of the same byte count."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tensor-fft_b200"))
import torch, torch.distributed as dist
import tfft
from tfft import dist as tdist

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << lg
m = n // world
x_re = torch.randn(m, device="cuda").to(torch.float16)
x_im = torch.randn(m, device="cuda").to(torch.float16)
mg = tdist.make_mg_plan(n)


def sync():
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()


def timed(fn, iters=5):
    fn(); sync()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)); dist.barrier()
    t = torch.tensor([sorted(ts)[len(ts) // 2]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

out = {"log2n": lg, "world": world}
mg.exec(x_re, x_im); sync()
for ph in range(3):
    out[f"phase{ph}_ms"] = round(timed(lambda: mg.exec_phase(ph, x_re, x_im)), 4)
out["exec_ms"] = round(timed(lambda: mg.exec(x_re, x_im)), 4)
def all_phases():
    for ph in range(3):
        mg.exec_phase(ph, x_re, x_im)
out["phases_back_to_back_no_barrier_ms"] = round(timed(all_phases), 4)
def insitu():
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    evs[0].record()
    for ph in range(3):
        mg.exec_phase(ph, x_re, x_im)
        evs[ph + 1].record()
    torch.cuda.synchronize()
    return [round(evs[i].elapsed_time(evs[i + 1]), 4) for i in range(3)]
sync(); insitu(); sync()
out["insitu_phase_ms"] = insitu()
sync()
def five0():
    for _ in range(5):
        mg.exec_phase(0, x_re, x_im)
out["exchange_x5_ms_each"] = round(timed(five0, iters=3) / 5, 4)
def ten():
    for _ in range(10):
        mg.exec(x_re, x_im)
out["exec_x10_ms_each"] = round(timed(ten, iters=3) / 10, 4)
# copy-engine peer copy and SM peer-store copy of the bytes one exchange sends to ONE peer... (world-1)/world of 4*m bytes
peer = (local + 1) % world
nbytes = 4 * m * (world - 1) // world
src = torch.empty(nbytes // 2, dtype=torch.float16, device=f"cuda:{local}")
dst = torch.empty(nbytes // 2, dtype=torch.float16, device=f"cuda:{peer}")
out["ce_peer_copy_ms"] = round(timed(lambda: dst.copy_(src, non_blocking=True)), 4)
out["ce_peer_copy_gbs"] = round(nbytes / out["ce_peer_copy_ms"] / 1e6, 1)
L = tfft.lib()
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
run = 1 << 20
cnt = (nbytes // 2) // run
def sm_copy():
    rc = L.tfft_copy_runs(src.data_ptr(), dst.data_ptr(), run, cnt, 1, 1, run, 0, 0, run, 0, 0, s)
    assert rc == 0, rc
out["sm_peer_store_ms"] = round(timed(sm_copy), 4)
out["sm_peer_store_gbs"] = round(nbytes / out["sm_peer_store_ms"] / 1e6, 1)
# cold remote pages: the same SM copy kernel cycling over 6 different remote destinations (and sources)
dsts = [torch.empty(nbytes // 2, dtype=torch.float16, device=f"cuda:{peer}") for _ in range(6)]
srcs = [torch.empty(nbytes // 2, dtype=torch.float16, device=f"cuda:{local}") for _ in range(6)]
for d_ in dsts:
    d_.copy_(src)      # enables peer access / touches the mapping once
sync()
def cyc():
    for i in range(6):
        rc = L.tfft_copy_runs(srcs[i].data_ptr(), dsts[i].data_ptr(), run, cnt, 1, 1, run, 0, 0, run, 0, 0, s)
        assert rc == 0, rc
out["sm_peer_store_cycling6_ms_each"] = round(timed(cyc, iters=3) / 6, 4)
dst2 = torch.empty_like(src)
def sm_copy_local():
    rc = L.tfft_copy_runs(src.data_ptr(), dst2.data_ptr(), run, cnt, 1, 1, run, 0, 0, run, 0, 0, s)
    assert rc == 0, rc
out["sm_local_copy_ms"] = round(timed(sm_copy_local), 4)
out["bytes_per_exchange_sent"] = nbytes
if rank == 0:
    print(json.dumps(out), flush=True)
sync()
mg.close()
dist.destroy_process_group()
