import os, sys
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/bench.py') else '.')
os.environ["PVACB_PROBE_VERBOSE"] = "1"
from pvac_hfhe_cppbyv_b200 import api
eng = api.Engine(0); eng.keygen(1)
print(eng.l2_gather_probe(3))
