import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import time_pair_score as t
t.case(4096, 256, 16, "bf16", "logit")
os.environ["MDG_FORCE_DIRECT_STORE"] = "1"
t.case(4096, 256, 16, "bf16", "logit")
