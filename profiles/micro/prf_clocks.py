"""Does the PRF kernel hold its clock? Samples nvidia-smi while prf_lpn_kernel runs back to back for a few seconds."""
import os, sys, subprocess, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pvac_hfhe_cppbyv_b200 import api
eng = api.Engine(0, prf_mode=api.PRF_FAITHFUL, tape=api.TAPE_SPLITMIX); eng.keygen(1)
rng = np.random.default_rng(1)
n = 4096
z, lo, hi = (rng.integers(0, 2**64, n, dtype=np.uint64) for _ in range(3))
rows = []; stop = False
def sampler():
    while not stop:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu", "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout.strip()
        rows.append((time.perf_counter(), o)); time.sleep(0.05)
th = threading.Thread(target=sampler); th.start()
eng.profile_enable(True)
t0 = time.perf_counter()
for rep in range(12):
    eng.stats_reset(); eng.profile_collect()
    eng.prf(z, lo, hi, 0)
    ms, _ = eng.profile_collect()["prf_lpn"]
    print(f"rep {rep}: t={time.perf_counter()-t0:.2f}s lpn {ms:.2f} ms {eng.stats()['aes_blocks']/ms/1e6:.2f} G blocks/s", flush=True)
stop = True; th.join()
for t, o in rows[::4]: print(f"{t-t0:6.2f}s {o}")
