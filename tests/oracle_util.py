"""ctypes binding of the CPU oracle (oracle/dfdally_oracle.cpp) -- test infrastructure only."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libdfdally_oracle.so")

EV_NAMES = ["KICKOFF", "REMOTE", "LOCAL", "ACK", "MN_BASE_NEW_MSG", "MN_BASE_SCHED_NEXT", "T_GENERATE", "T_ARRIVE",
            "T_SEND", "T_BUFFER", "R_SEND", "R_ARRIVE", "R_BUFFER", "R_SNAPSHOT"]


class Result(ctypes.Structure):
    _fields_ = [("events", ctypes.c_ulonglong), ("ev_by_type", ctypes.c_ulonglong * 16),
                ("total_hops", ctypes.c_longlong), ("finished_packets", ctypes.c_longlong),
                ("finished_msgs", ctypes.c_longlong), ("finished_chunks", ctypes.c_longlong),
                ("total_msg_sz", ctypes.c_longlong),
                ("total_time", ctypes.c_double), ("max_latency", ctypes.c_double),
                ("packet_gen", ctypes.c_longlong), ("packet_fin", ctypes.c_longlong),
                ("local_sr", ctypes.c_longlong), ("local_sg", ctypes.c_longlong), ("remote", ctypes.c_longlong),
                ("minimal_count", ctypes.c_longlong), ("nonmin_count", ctypes.c_longlong),
                ("svr_msgs_recvd", ctypes.c_longlong),
                ("svr_sum_latency", ctypes.c_double), ("svr_max_latency", ctypes.c_double),
                ("max_end_ts", ctypes.c_double), ("last_event_ts", ctypes.c_double),
                ("rng_draws", ctypes.c_ulonglong), ("run_seconds", ctypes.c_double),
                ("error", ctypes.c_char * 512)]


class OracleError(RuntimeError):
    pass


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


_lib = None


def load():
    global _lib
    if _lib is None:
        src = os.path.join(ORACLE_DIR, "dfdally_oracle.cpp")
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            build()
        lib = ctypes.CDLL(LIB)
        lib.dfdally_oracle_run.restype = ctypes.c_int
        lib.dfdally_oracle_run.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                           ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_char_p,
                                           ctypes.c_int, ctypes.POINTER(Result)]
        lib.dfdally_oracle_rng_stream.restype = None
        lib.dfdally_oracle_rng_stream.argtypes = [ctypes.c_ulonglong, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                                  ctypes.POINTER(ctypes.c_longlong)]
        lib.dfdally_oracle_bytes_to_ns.restype = ctypes.c_double
        lib.dfdally_oracle_bytes_to_ns.argtypes = [ctypes.c_ulonglong, ctypes.c_double]
        _lib = lib
    return _lib


def run(conf, lp_io_dir=None, traffic=1, num_messages=20, payload_sz=2048, arrival_time=0.0, load_per_svr=0.0,
        warm_up_time=0.0, end_time=0.0):
    lib = load()
    res = Result()
    rc = lib.dfdally_oracle_run(os.fsencode(conf), traffic, num_messages, payload_sz, arrival_time, load_per_svr,
                                warm_up_time, end_time, os.fsencode(lp_io_dir) if lp_io_dir else None, 1,
                                ctypes.byref(res))
    if rc != 0:
        raise OracleError(res.error.decode(errors="replace"))
    return res


def rng_stream(stream, n):
    lib = load()
    out = (ctypes.c_double * n)()
    st = (ctypes.c_longlong * 4)()
    lib.dfdally_oracle_rng_stream(stream, n, out, st)
    return list(out), list(st)


LPIO_FILES = ["dragonfly-cn-stats", "dragonfly-link-stats", "model-net-category-all", "model-net-category-test",
              "synthetic-stats", "raw-lp-stats", "summary.txt"]


OPTIONAL_FILES = ["dragonfly-snapshots.csv"]  # only with PARAMS router_buffer_snapshots


def read_dir(d):
    out = {}
    for f in LPIO_FILES + OPTIONAL_FILES:
        p = os.path.join(d, f)
        if os.path.exists(p):
            with open(p, "rb") as fh:
                out[f] = fh.read()
    return out


# ---- reference-source oracle: the reference's own sources compiled unchanged against oracle/ref/shim -> oracle/_ref
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "model-net-synthetic-dragonfly-all")
REF_FILES = ["dragonfly-snapshots.csv", "dragonfly-cn-stats", "dragonfly-link-stats", "model-net-category-all", "model-net-category-test",
             "synthetic-stats"]


def ref_available():
    return os.path.exists(REF_BIN)


def build_ref():
    """Only possible where /root/reference exists (the authoring container)."""
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", os.path.join(ORACLE_DIR, "ref")], check=True)
    return ref_available()


def run_ref(conf, lp_io_dir, traffic=1, num_messages=20, payload_sz=2048, arrival_time=0.0, load_per_svr=0.0,
            warm_up_time=0.0, end_time=0.0):
    """Run the reference binary twin (sequential shim engine). Returns (events, event-loop seconds, stdout)."""
    import re
    import shutil
    conf = os.path.abspath(conf)
    shutil.rmtree(lp_io_dir, ignore_errors=True)  # lp_io_prepare refuses an existing directory
    args = [REF_BIN, "--sync=1", f"--traffic={traffic}", f"--num_messages={num_messages}", f"--payload_sz={payload_sz}"]
    if arrival_time:
        args.append(f"--arrival_time={arrival_time}")
    if load_per_svr:
        args.append(f"--load_per_svr={load_per_svr}")
    if warm_up_time:
        args.append(f"--warm_up_time={warm_up_time}")
    if end_time:
        args.append(f"--end={end_time}")
    args += [f"--lp-io-dir={os.path.abspath(lp_io_dir)}", "--", conf]
    r = subprocess.run(args, cwd=os.path.dirname(conf), capture_output=True, text=True)  # link files are cwd-relative
    if r.returncode != 0:
        raise OracleError("reference-source oracle failed: " + r.stderr[-500:])
    ev = int(re.search(r"Net Events Processed\s+(\d+)", r.stdout).group(1))
    m = re.search(r"shim: (\d+) events in ([0-9.]+) s", r.stderr)
    return ev, float(m.group(2)), r.stdout


def stdout_stats_block(text):
    """The statistics the driver prints after the run (dally:3169-3190, synthetic:635-657), without blank lines."""
    lines = text.splitlines()
    for i, l in enumerate(lines):
        if l.startswith("Average number of hops"):
            return [x for x in lines[i:] if x.strip()]
    return []
