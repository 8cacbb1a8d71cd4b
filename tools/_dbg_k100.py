import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "_r1") if os.environ.get("USE_R1") else ROOT)
import local_rag_system_b200 as rag
print(rag.__file__)
rows, dim, k, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
kw = {} if os.environ.get("USE_R1") else {"rerank": False}
st = rag.DeviceStore(dim, "bf16", "cosine", capacity_hint=rows, **kw)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
for s in range(0, rows, 500_000):
    m = min(500_000, rows - s)
    x = torch.nn.functional.normalize(torch.randn((m, dim), generator=gen, device="cuda"), dim=1)
    torch.cuda.synchronize(); st.upsert_device(x.data_ptr(), m)
rng = np.random.default_rng(0)
for it in range(4):
    q = rng.standard_normal((B, dim), dtype=np.float32)
    r, d, c = st.query(q, k, regime="tensor")
    print("ok", rows, dim, k, B, round(st.last_query_info()["kernel_ms"], 2), r[0, :3], flush=True)
