#!/usr/bin/env python
"""The flow of the reference's driver script /root/reference/tests/correlations.py (:29-110) on this
library, line for line where a line exists: read / make a catalogue, CIC-paint it, density contrast,
P(k) multipoles in the script's bins, xi(s), and the bispectrum call of tests/bispec.py:53-56.
Needs a B200 (there is no CPU fallback).  Plotting and the Pylians3 cross-checks of the script are left out.

    python examples/correlations_flow.py [catalogue.dat]

Without a file (the reference's Patchy mock is not distributed with it) a lognormal mock of the same
size is generated on the device with the reference's own recipe (tests/create_lognormal.py:44-55).
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from jax_powspec_b200 import read_catalog_text                                     # np.loadtxt + box mask
from jax_powspec_b200.correlations import bispec, powspec_vec, xi_vec              # was: from src.correlations import ...
from jax_powspec_b200.mas import cic_mas_vec                                       # was: from src.mas import ...
from jax_powspec_b200.mocks import lognormal_mock

n_bins = 256
box_size = 2500.0
k_ny = np.pi * n_bins / box_size

if len(sys.argv) > 1:
    # tests/correlations.py:29-31: loadtxt(usecols=(0,1,2)), keep 0 < x,y,z < box, device_put
    particles = read_catalog_text(sys.argv[1], usecols=(0, 1, 2), box_size=box_size)
else:
    klin = np.linspace(1e-4, 10, 4056)
    plin = 2.0e4 * (klin / 0.02) / (1.0 + (klin / 0.02) ** 2) ** 1.7               # stand-in for jax_cosmo's linear P(k)
    particles = lognormal_mock(n_bins, klin, plin, 1.1, 3.2e-4, 1005638091, box_size)
n_part = particles.shape[0]
print(n_part)
shot_noise = box_size ** 3 / n_part

w = torch.ones(n_part, device=particles.device)
torch.cuda.synchronize()
s = time.time()
delta = torch.zeros((n_bins, n_bins, n_bins), device=particles.device)
delta = cic_mas_vec(delta, particles[:, 0], particles[:, 1], particles[:, 2], w, n_part,
                    float(particles[:, 0].min()), float(particles[:, 1].min()), float(particles[:, 2].min()),
                    box_size, n_bins, True)
delta /= delta.mean()
delta -= 1.0
torch.cuda.synchronize()
print(f"MAS took {time.time() - s} s.", flush=True)

kedges = np.arange(1e-4, 5, 0.2e-2)                                                # tests/correlations.py:76
k, pk, modes = powspec_vec(delta, box_size, kedges)
mask = k < k_ny
k, pk = k[mask], pk[mask]
print("P0(k) - shot noise, first non-empty bins:", (pk[:, 0] - shot_noise)[~torch.isnan(pk[:, 0])][:5].tolist())

s_edges = np.linspace(0, 200, 41)                                                  # tests/correlations.py:92
r, xi, modes = xi_vec(delta, box_size, s_edges)
print("r^2 xi0(r), bins 1..5:", (r ** 2 * xi[:, 0])[1:6].tolist())

theta = np.linspace(0, np.pi, 20)                                                  # tests/bispec.py:53-54
s = time.time()
k_all, pk_shell, theta, B, Q = bispec(delta, box_size, 0.1, 0.2, theta)
torch.cuda.synchronize()
print(f"bispectrum took {time.time() - s} s.  Q(theta):", Q[:5].tolist())
