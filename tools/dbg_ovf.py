import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_b200 import capi
dev = torch.device("cuda", 0)
n, d, K, nq = 1_000_000, 768, 100, 512
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.randn((n, d), generator=g, device=dev)
ix = capi.FlatIndex(d, capi.COSINE)
ix.add_device(np.arange(1, n + 1, dtype=np.uint32), x.data_ptr(), n)
del x
q = torch.randn((nq, d), generator=g, device=dev)
oi = torch.zeros((nq, K), dtype=torch.int32, device=dev); osc = torch.zeros((nq, K), device=dev); oc = torch.zeros(nq, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
os.environ["COMET_B200_DBG_STAGED"] = "1"
for rep in range(3):
    ix.search_device(q.data_ptr(), nq, K, oi.data_ptr(), osc.data_ptr(), oc.data_ptr(), K, stream=st, path=capi.PATH_TENSOR)
    torch.cuda.synchronize()
    c = oc.cpu().numpy()
    print("rep", rep, "overflowed", int((c < 0).sum()), np.nonzero(c < 0)[0][:10], ix.last_stats(), flush=True)
