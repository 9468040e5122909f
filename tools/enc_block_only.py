import os, sys, json
ROOT = os.getcwd(); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, bench
dev = torch.device("cuda:0")
e = bench.encoder_block(dev)
print(os.environ.get("MDG_TASKS_PER_CTA", "default"), " ".join("%s %.2f ms" % (k.split("_")[-1], v["ms"]) for k, v in e.items()))
