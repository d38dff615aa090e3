"""rdp-validate-dump equivalent (reference: rdp_validate_dump.cpp:84-198): replay an RDPDUMP2 file through
RDP::CommandProcessor and through the Angrylion oracle, comparing RDRAM + hidden RDRAM at every SIGNAL_COMPLETE and the
scan-out at every END_FRAME.

    python tests/validate_dump.py dump.rdp [--backend cuda|hostsim]

Lives under tests/ because it drives the oracle (test infrastructure)."""
from __future__ import annotations

import sys

import numpy as np

from sidebyside import SideBySide, build_hostsim, rdp


def validate_dump(backend, path, lib=None, log=None):
    """Replay through the processor with rdp_b200_dump_replay and through the oracle record by record; compare RDRAM and
    hidden RDRAM at every SIGNAL_COMPLETE and the scan-out at every END_FRAME."""
    probe = rdp.DumpPlayer(path, library=lib)
    rdram_size, hidden_size = probe.get_rdram_size(), probe.get_hidden_rdram_size()
    probe.close()
    sbs = SideBySide(backend, rdram_size=rdram_size, hidden_size=hidden_size)
    a = rdp.DumpPlayer(path, library=lib)  # drives the processor natively
    b = rdp.DumpPlayer(path, library=lib)  # drives the oracle from Python
    errs, checks = [], 0
    try:
        while True:
            n, stop = a.replay(sbs.gpu)
            assert n >= 0
            while True:
                kind, x, y, payload = b.next_event()
                if kind == rdp.DUMP_UPDATE_RDRAM:
                    sbs.oracle.rdram[:] = payload
                elif kind == rdp.DUMP_UPDATE_HIDDEN_RDRAM:
                    sbs.oracle.hidden[:hidden_size] = payload
                elif kind == rdp.DUMP_RDP_COMMAND:
                    sbs.oracle.command_stream(payload)
                elif kind == rdp.DUMP_SET_VI_REGISTER:
                    sbs.oracle.set_vi_register(x, y)
                else:
                    break
            assert kind == stop
            if stop == rdp.DUMP_EOF:
                break
            if stop == rdp.DUMP_SIGNAL_COMPLETE:
                errs += sbs.compare_rdram()
            else:
                ref, pix, w, h = sbs.end_frame()
                if ref is None or ref.shape != pix.shape or not np.array_equal(ref & 0xFFFFFF, pix & 0xFFFFFF):
                    errs.append("scan-out differs")
            checks += 1
            if log:
                log(f"passed check {checks} ({'sync' if stop == rdp.DUMP_SIGNAL_COMPLETE else 'frame'})")
            if errs:
                break
    finally:
        a.close()
        b.close()
        sbs.close()
    return errs, checks




if __name__ == "__main__":
    backend = "cuda"
    if "--backend" in sys.argv:
        backend = sys.argv[sys.argv.index("--backend") + 1]
    library = build_hostsim() if backend == "hostsim" else None
    errors, n = validate_dump(backend, sys.argv[1], library, log=print)
    for e in errors:
        print("FAILED:", e)
    print(f"{n} checks, {'FAILED' if errors else 'all passed'}")
    sys.exit(1 if errors else 0)
