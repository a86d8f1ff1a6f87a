import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_b200 import capi
L = capi.lib()
dev = torch.device("cuda", 0)
n, d, K, nq = 1_000_000, 768, 100, 512
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.randn((n, d), generator=g, device=dev)
ix = capi.FlatIndex(d, capi.METRICS[os.environ.get("DBG_METRIC", "cosine")])
ix.add_device(np.arange(1, n + 1, dtype=np.uint32), x.data_ptr(), n)
del x
q = torch.randn((nq, d), generator=g, device=dev)
oi = torch.zeros((nq, K), dtype=torch.int32, device=dev); osc = torch.zeros((nq, K), device=dev); oc = torch.zeros(nq, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
def run(tag, reps=10):
    for _ in range(3):
        ix.search_device(q.data_ptr(), nq, K, oi.data_ptr(), osc.data_ptr(), oc.data_ptr(), K, stream=st, path=capi.PATH_TENSOR)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ix.search_device(q.data_ptr(), nq, K, oi.data_ptr(), osc.data_ptr(), oc.data_ptr(), K, stream=st, path=capi.PATH_TENSOR)
    e1.record(); torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / reps
    L.cm_profile_reset(); L.cm_profile_enable(1)
    for _ in range(5):
        ix.search_device(q.data_ptr(), nq, K, oi.data_ptr(), osc.data_ptr(), oc.data_ptr(), K, stream=st, path=capi.PATH_TENSOR)
    torch.cuda.synchronize(); L.cm_profile_enable(0)
    gm, gn = capi.profile_get(capi.PROF_FLAT_GEMM); sm, sn = capi.profile_get(capi.PROF_SELECT); rm, rn = capi.profile_get(capi.PROF_RESCORE)
    print(f"{tag}: step {total:.3f} ms | gemm {gm/5:.3f} ({gn//5} launches) select {sm/5:.3f} rescore {rm/5:.3f} min cnt {int(oc.min())}", flush=True)
for arg in sys.argv[1:]:
    for kv in arg.split(","):
        k, v = kv.split("=")
        os.environ[k] = v
    run(arg)
