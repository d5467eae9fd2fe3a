"""Debug: device batch builder (static mode) vs host collate on the bench's builder dataset; plan meta after a step."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feta_tmlr_b200 import synthetic, data as fdata, engine, ops
name = sys.argv[1] if len(sys.argv) > 1 else "ZINC"
cfg = synthetic.CONFIGS[name]; B = cfg['batch']
dev = torch.device("cuda")
pool = synthetic.make_dataset(name, B * 4, seed=100)
pstore = fdata.GraphStore(pool, kind=cfg['kind'], n_tags=cfg['n_tags'])
caps = engine.static_caps(pstore, B)
print("caps", caps)
graphs = synthetic.make_dataset(name, B * 16, seed=0)
store = fdata.GraphStore(graphs, kind=cfg['kind'], n_tags=cfg['n_tags'])
builder = fdata.DeviceBatchBuilder(store, dev)
bad = 0
for i in range(16):
    ids = np.arange(i * B, (i + 1) * B)
    try:
        hb = fdata.collate_host(store, ids, static=caps)
    except ValueError as e:
        print(i, "host ValueError", e); continue
    db = builder.build(ids, static=caps)
    torch.cuda.synchronize()
    for k, (h, d_) in enumerate(zip(hb[:7], db[:7])):
        if h is None: continue
        h = torch.as_tensor(h)
        if not torch.equal(h.to(d_.dtype), d_.cpu()):
            bad += 1
            print("batch", i, "field", k, "differs", h.shape, d_.shape, (h.to(d_.dtype) != d_.cpu()).sum().item())
print("mismatching fields:", bad)
