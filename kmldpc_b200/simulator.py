"""Mirror of the reference's `Simulator` (kmldpc/include/simulator.h:56-60, src/simulator.cc:3-67) for Python callers:
construct from a config.toml, call simulate(), get the BER/FER tables — same keys, same table format, GPU path."""
from __future__ import annotations

import ctypes as C
import os
import sys
import time

import numpy as np

from . import capi
from .link import CONFIG_DIR, KmlError


class Simulator:
    def __init__(self, config_toml: str | None = None, data_dir: str | None = None, **overrides):
        self._lib = capi.load()
        self.cfg = capi.KmlSweepCfg()
        path = config_toml or os.path.join(CONFIG_DIR, "config.toml")
        rc = self._lib.kml_sweep_cfg_load(path.encode(), C.byref(self.cfg))
        if rc != 0:
            raise KmlError(f"config: {self._lib.kml_last_error(None).decode()}")
        self.data_dir = data_dir or os.path.dirname(os.path.abspath(path))
        for k, v in overrides.items():
            if k in ("matrix_file", "modem_file"):
                v = v.encode()
            setattr(self.cfg, k, v)
        self.lines: list[str] = []

    def last_timing(self) -> tuple[float, float]:
        """(setup seconds, sweep seconds) of the last kml_sweep_run of this process: code / constellation loading, context
        creation and NCCL initialisation, then the SNR points themselves."""
        t = (C.c_double * 2)()
        self._lib.kml_sweep_last_timing(t)
        return float(t[0]), float(t[1])

    @property
    def n_points(self) -> int:
        return self._lib.kml_sweep_points(C.byref(self.cfg))

    def simulate(self, echo: bool = True):
        n = self.n_points
        ber = np.zeros(n, np.float64)
        fer = np.zeros(n, np.float64)
        cnt = np.zeros((n, 4), np.uint64)

        def on_line(line, _user):
            s = line.decode()
            self.lines.append(s)
            if echo:  # the reference's logger prefix (lib/lab/src/log.cc:90-100)
                sys.stdout.write(f"[{time.strftime('%Y-%m-%d %H:%M:%S')}] \x1b[32;1m[INFO]\x1b[0m {s}\n")

        cb = capi.LOG_CB(on_line)
        rc = self._lib.kml_sweep_run(C.byref(self.cfg), self.data_dir.encode(),
                                     ber.ctypes.data_as(capi.c_f64p), fer.ctypes.data_as(capi.c_f64p),
                                     cnt.ctypes.data_as(capi.c_u64p), cb, None)
        if rc != 0:
            raise KmlError(f"sweep: {self._lib.kml_last_error(None).decode()} (rc={rc})")
        snr = self.cfg.min_snr + self.cfg.step_snr * np.arange(n)
        return snr, ber, fer, cnt
