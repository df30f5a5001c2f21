"""Two GPUs: export rate of ONE GPU's product batch, direct and relayed through the other GPU (the other GPU idle). If the peer hop is
a real NVLink copy the relayed rate is that of the relay GPU's host link; if the driver stages it through host memory it is about half."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from pvac_hfhe_cppbyv_b200 import api
eng = api.Engine(0, prf_mode=api.PRF_LIVE, tape=api.TAPE_SPLITMIX); eng.keygen(1)
v = np.arange(512, dtype=np.uint64)
P = eng.ct_mul(eng.enc_value(v, 1), eng.enc_value(v + 3, 2), 3)
n, nl, ne, by = eng.blob_info(P)
h = torch.empty(by, dtype=torch.uint8).pin_memory(); hv = h.numpy()
ref = None
for relay in (-1, 1, -1, 1):
    eng.set_export_relay(relay)
    for _ in range(2):
        eng.export_blob_async(P, hv); eng.export_wait()
    t = time.perf_counter()
    for _ in range(6):
        eng.export_blob_async(P, hv)
    eng.export_wait()
    dt = time.perf_counter() - t
    if ref is None: ref = hv.copy()
    print(f"relay {relay}: {6 * by / dt / 1e9:.1f} GB/s ({by / 1e6:.0f} MB per export), bytes equal: {bool((hv == ref).all())}")
