"""codes_b200 -- B200-native engine for ONE hot path of CODES: the dragonfly-dally network model driven by
the synthetic traffic generator.

This package is a thin ctypes binding over the C-ABI library `libdfdally_b200.so`
(`include/dfdally_b200.h`).  The class and function names mirror the reference's call sequence
(`src/network-workloads/model-net-synthetic-dragonfly-all.c:660-817`): configuration_load ->
model_net_register -> codes_mapping_setup -> model_net_configure -> tw_run -> lp_io_flush ->
model_net_report_stats.  There is no CPU fallback: without the built CUDA library or without a GPU every
compute call raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdfdally_b200.so")

# enum TRAFFIC, model-net-synthetic-dragonfly-all.c:79-87
UNIFORM, RAND_PERM, NEAREST_GROUP, NEAREST_NEIGHBOR, RANDOM_OTHER_GROUP = 1, 2, 3, 4, 5

EVENT_TYPE_NAMES = ["KICKOFF", "REMOTE", "LOCAL", "ACK", "MN_BASE_NEW_MSG", "MN_BASE_SCHED_NEXT", "T_GENERATE",
                    "T_ARRIVE", "T_SEND", "T_BUFFER", "R_SEND", "R_ARRIVE", "R_BUFFER"]


class SyntheticOpts(ctypes.Structure):
    """tw_optdef app_opt[] of the synthetic driver (synthetic:173-194)."""
    _fields_ = [("traffic", ctypes.c_int), ("num_messages", ctypes.c_int), ("payload_sz", ctypes.c_int),
                ("rperm_threshold", ctypes.c_longlong), ("arrival_time", ctypes.c_double),
                ("load_per_svr", ctypes.c_double), ("warm_up_time", ctypes.c_double), ("end_time", ctypes.c_double)]


class Summary(ctypes.Structure):
    _fields_ = [("events", ctypes.c_ulonglong), ("events_by_type", ctypes.c_ulonglong * 16),
                ("windows", ctypes.c_ulonglong),
                ("total_hops", ctypes.c_longlong), ("finished_packets", ctypes.c_longlong),
                ("finished_msgs", ctypes.c_longlong), ("finished_chunks", ctypes.c_longlong),
                ("total_msg_sz", ctypes.c_longlong),
                ("total_time", ctypes.c_double), ("max_latency", ctypes.c_double),
                ("packet_gen", ctypes.c_longlong), ("packet_fin", ctypes.c_longlong),
                ("local_same_router", ctypes.c_longlong), ("local_same_group", ctypes.c_longlong),
                ("remote", ctypes.c_longlong),
                ("minimal_count", ctypes.c_longlong), ("nonmin_count", ctypes.c_longlong),
                ("svr_msgs_recvd", ctypes.c_longlong),
                ("svr_sum_latency", ctypes.c_double), ("svr_max_latency", ctypes.c_double),
                ("max_end_ts", ctypes.c_double),
                ("total_offered_load", ctypes.c_double), ("total_observed_load", ctypes.c_double),
                ("rng_draws", ctypes.c_ulonglong), ("window_ns", ctypes.c_double),
                ("h2d_bytes", ctypes.c_ulonglong), ("d2h_bytes", ctypes.c_ulonglong),
                ("device_bytes", ctypes.c_ulonglong),
                ("grid_blocks", ctypes.c_int), ("block_threads", ctypes.c_int), ("num_sms", ctypes.c_int),
                ("num_terminals", ctypes.c_int), ("num_routers", ctypes.c_int), ("radix", ctypes.c_int),
                ("cyc_busy", ctypes.c_double), ("cyc_idle", ctypes.c_double), ("cyc_busy_max_warp", ctypes.c_double),
                ("activations", ctypes.c_ulonglong), ("passes", ctypes.c_ulonglong)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if name == "events_by_type" else v
        return d


class EngineError(RuntimeError):
    pass


_lib = None

# every symbol include/dfdally_b200.h declares: (name, restype, argtypes)
_P = ctypes.c_void_p
_S = ctypes.c_char_p
_I = ctypes.c_int
_LL = ctypes.c_longlong
ABI = [
    ("dfd_synthetic_opts_default", None, [ctypes.POINTER(SyntheticOpts)]),
    ("dfd_engine_create", _I, [ctypes.POINTER(_P)]),
    ("dfd_engine_destroy", None, [_P]),
    ("dfd_last_error", _S, [_P]),
    ("dfd_configuration_load", _I, [_P, _S]),
    ("dfd_configuration_get_value", _I, [_P, _S, _S, _S, _I]),
    ("dfd_model_net_register", _I, [_P]),
    ("dfd_lp_type_lookup", _I, [_P, _S]),
    ("dfd_codes_mapping_setup", _I, [_P]),
    ("dfd_codes_mapping_get_lp_count", _LL, [_P, _S]),
    ("dfd_codes_mapping_get_lpid_from_relative", _LL, [_P, _LL, _S]),
    ("dfd_codes_mapping_get_lp_relative_id", _LL, [_P, _LL]),
    ("dfd_model_net_configure", _I, [_P, ctypes.POINTER(_I), _I, ctypes.POINTER(_I)]),
    ("dfd_synthetic_configure", _I, [_P, ctypes.POINTER(SyntheticOpts)]),
    ("dfd_engine_upload", _I, [_P, _P]),
    ("dfd_engine_reset", _I, [_P, _P]),
    ("dfd_engine_run", _I, [_P, _P]),
    ("dfd_engine_finish", _I, [_P, _P]),
    ("dfd_tw_run", _I, [_P, _P]),
    ("dfd_engine_set_ensemble", _I, [_P, _I]),
    ("dfd_abi_sizes", None, [_P, _P]),
    ("dfd_engine_set_partition", _I, [_P, _I, _I]),
    ("dfd_engine_ipc_export", _I, [_P, _P]),
    ("dfd_engine_ipc_import", _I, [_P, _P, _I]),
    ("dfd_engine_ipc_ready", _I, [_P]),
    ("dfd_engine_attach_local_peers", _I, [_P, ctypes.POINTER(_P), _I]),
    ("dfd_engine_partition_bytes", _LL, [_P]),
    ("dfd_partition_bounds", _I, [_I, _I, _I, _I, _I, ctypes.POINTER(_I)]),
    ("dfd_engine_export_partition", _I, [_P, _P, _LL]),
    ("dfd_engine_import_partition", _I, [_P, _P, _LL]),
    ("dfd_engine_finalize", _I, [_P]),
    ("dfd_lp_io_flush", _I, [_P, _S]),
    ("dfd_model_net_report_stats", _I, [_P, _S, _I]),
    ("dfd_get_summary", _I, [_P, ctypes.POINTER(Summary)]),
    ("dfd_run_synthetic", _I, [_S, ctypes.POINTER(SyntheticOpts), _S, ctypes.POINTER(Summary), _S, _I]),
    ("dfd_bytes_to_ns", ctypes.c_double, [ctypes.c_ulonglong, ctypes.c_double]),
    ("dfd_cuda_device_count", _I, []),
]


def bind_library(path):
    """ctypes-bind one build of the C-ABI (every symbol of include/dfdally_b200.h) and check the struct mirrors."""
    lib = ctypes.CDLL(path)
    for name, res, args in ABI:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    # the two structs that cross the boundary by pointer must mirror this build of the library exactly
    a, b = ctypes.c_int(0), ctypes.c_int(0)
    lib.dfd_abi_sizes(ctypes.byref(a), ctypes.byref(b))
    if (a.value, b.value) != (ctypes.sizeof(Summary), ctypes.sizeof(SyntheticOpts)):
        raise EngineError(f"{path}: dfd_summary / dfd_synthetic_opts are {a.value} / {b.value} bytes in the library but "
                          f"{ctypes.sizeof(Summary)} / {ctypes.sizeof(SyntheticOpts)} in this binding: rebuild one of them")
    return lib


def load_library():
    """Load libdfdally_b200.so (built by __graft_entry__.build() / `make -C codes_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(f"{LIB_PATH} is missing: build it with `make -C codes_b200/csrc` "
                          "(there is no CPU fallback for the dragonfly-dally engine)")
    _lib = bind_library(LIB_PATH)
    return _lib


def bytes_to_ns(nbytes, gib_per_s):
    """dragonfly-dally.cxx:1588-1599."""
    return load_library().dfd_bytes_to_ns(nbytes, gib_per_s)


def cuda_device_count():
    return load_library().dfd_cuda_device_count()


def partition_bounds(num_groups, routers_per_group, cns_per_router, rank, world):
    """(first node, end node, first router, end router) owned by `rank` of `world` (contiguous groups)."""
    out = (_I * 4)()
    if load_library().dfd_partition_bounds(num_groups, routers_per_group, cns_per_router, rank, world, out) != 0:
        raise EngineError(f"bad partition: rank {rank} of {world} over {num_groups} groups")
    return tuple(out)


def make_opts(traffic=UNIFORM, num_messages=20, payload_sz=2048, rperm_threshold=131072, arrival_time=0.0,
              load_per_svr=0.0, warm_up_time=0.0, end_time=0.0):
    return SyntheticOpts(traffic, num_messages, payload_sz, rperm_threshold, arrival_time, load_per_svr,
                         warm_up_time, end_time)


class Engine:
    """One simulation context (the reference keeps this in file-scope globals)."""

    def __init__(self, lib=None):
        self._lib = lib or load_library()
        self._h = _P()
        if self._lib.dfd_engine_create(ctypes.byref(self._h)) != 0 or not self._h:
            raise EngineError("dfd_engine_create failed")

    def close(self):
        if self._h:
            self._lib.dfd_engine_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self._lib.dfd_last_error(self._h).decode(errors="replace"))

    # --- the reference's call sequence ---------------------------------------------------------------
    def configuration_load(self, conf_path):
        self._ck(self._lib.dfd_configuration_load(self._h, os.fsencode(conf_path)))

    def configuration_get_value(self, section, key):
        buf = ctypes.create_string_buffer(1024)
        n = self._lib.dfd_configuration_get_value(self._h, section.encode(), key.encode(), buf, 1024)
        return buf.value.decode() if n > 0 else None

    def model_net_register(self):
        self._ck(self._lib.dfd_model_net_register(self._h))

    def lp_type_lookup(self, name):
        return bool(self._lib.dfd_lp_type_lookup(self._h, name.encode()))

    def codes_mapping_setup(self):
        self._ck(self._lib.dfd_codes_mapping_setup(self._h))

    def codes_mapping_get_lp_count(self, lp_type_name):
        return self._lib.dfd_codes_mapping_get_lp_count(self._h, lp_type_name.encode())

    def codes_mapping_get_lpid_from_relative(self, rel, lp_type_name):
        return self._lib.dfd_codes_mapping_get_lpid_from_relative(self._h, rel, lp_type_name.encode())

    def codes_mapping_get_lp_relative_id(self, gid):
        return self._lib.dfd_codes_mapping_get_lp_relative_id(self._h, gid)

    def model_net_configure(self):
        ids = (_I * 4)()
        n = _I(0)
        self._ck(self._lib.dfd_model_net_configure(self._h, ids, 4, ctypes.byref(n)))
        return list(ids[:n.value])

    def synthetic_configure(self, opts):
        self._ck(self._lib.dfd_synthetic_configure(self._h, ctypes.byref(opts)))

    def setup(self, conf_path, opts):
        """configuration_load .. synthetic_configure in the driver's order."""
        self.configuration_load(conf_path)
        self.model_net_register()
        self.codes_mapping_setup()
        ids = self.model_net_configure()
        self.synthetic_configure(opts)
        return ids

    # --- tw_run() ------------------------------------------------------------------------------------
    def upload(self, stream=None):
        self._ck(self._lib.dfd_engine_upload(self._h, _P(stream or 0)))

    def reset(self, stream=None):
        self._ck(self._lib.dfd_engine_reset(self._h, _P(stream or 0)))

    def run(self, stream=None):
        self._ck(self._lib.dfd_engine_run(self._h, _P(stream or 0)))

    def finish(self, stream=None):
        self._ck(self._lib.dfd_engine_finish(self._h, _P(stream or 0)))

    def tw_run(self, stream=None):
        self._ck(self._lib.dfd_tw_run(self._h, _P(stream or 0)))

    def set_ensemble(self, simulations):
        """This engine will share the GPU with simulations-1 others (call before upload)."""
        self._ck(self._lib.dfd_engine_set_ensemble(self._h, int(simulations)))

    # --- group-sharded runs (one process per GPU) ----------------------------------------------------
    def set_partition(self, rank, world):
        self._ck(self._lib.dfd_engine_set_partition(self._h, rank, world))

    def ipc_export(self):
        buf = ctypes.create_string_buffer(64)
        self._ck(self._lib.dfd_engine_ipc_export(self._h, buf))
        return buf.raw

    def ipc_import(self, handles):
        blob = b"".join(handles)
        self._ck(self._lib.dfd_engine_ipc_import(self._h, blob, len(handles)))

    def ipc_ready(self):
        return bool(self._lib.dfd_engine_ipc_ready(self._h))

    def export_partition(self):
        n = self._lib.dfd_engine_partition_bytes(self._h)
        buf = ctypes.create_string_buffer(n)
        self._ck(self._lib.dfd_engine_export_partition(self._h, buf, n))
        return buf.raw

    def import_partition(self, blob):
        self._ck(self._lib.dfd_engine_import_partition(self._h, blob, len(blob)))

    def finalize(self):
        self._ck(self._lib.dfd_engine_finalize(self._h))

    # --- outputs -------------------------------------------------------------------------------------
    def lp_io_flush(self, directory):
        self._ck(self._lib.dfd_lp_io_flush(self._h, os.fsencode(directory)))

    def model_net_report_stats(self):
        buf = ctypes.create_string_buffer(16384)
        self._ck(self._lib.dfd_model_net_report_stats(self._h, buf, 16384))
        return buf.value.decode()

    def summary(self):
        s = Summary()
        self._lib.dfd_get_summary(self._h, ctypes.byref(s))
        return s


def run_synthetic(conf_path, lp_io_dir=None, lib=None, **opts):
    """The whole main() of model-net-synthetic-dragonfly-all on the GPU; returns a Summary."""
    lib = lib or load_library()
    o = make_opts(**opts)
    s = Summary()
    err = ctypes.create_string_buffer(1024)
    rc = lib.dfd_run_synthetic(os.fsencode(conf_path), ctypes.byref(o), os.fsencode(lp_io_dir) if lp_io_dir else None,
                               ctypes.byref(s), err, 1024)
    if rc != 0:
        raise EngineError(err.value.decode(errors="replace"))
    return s


def run_ensemble(conf_path, variants, lp_io_dirs=None, streams=None):
    """Several independent simulations of one network at the same time on ONE GPU (a parameter sweep): `variants` is
    a list of option dicts (as for run_synthetic), every simulation gets its own engine and CUDA stream, the event
    loops run concurrently.  Returns the list of Summaries; results are identical to running each variant alone."""
    import torch
    k = len(variants)
    streams = streams or [torch.cuda.Stream() for _ in range(k)]
    engines = []
    try:
        for v, s in zip(variants, streams):
            e = Engine()
            engines.append(e)
            e.setup(conf_path, make_opts(**v))
            e.set_ensemble(k)
            e.upload(s.cuda_stream)
        for e, s in zip(engines, streams):
            e.run(s.cuda_stream)
        out = []
        for i, (e, s) in enumerate(zip(engines, streams)):
            e.finish(s.cuda_stream)
            if lp_io_dirs and lp_io_dirs[i]:
                e.lp_io_flush(lp_io_dirs[i])
            out.append(e.summary())
        return out
    finally:
        for e in engines:
            e.close()


def sharded_tw_run(eng, dist, stream=None):
    """tw_run() of a group-sharded engine (set_partition done): upload, exchange of the CUDA-IPC handles when the device
    arrays are new, event loop, download, merge of the partitions on rank 0.  Collective over `dist`; a failure on any
    rank is raised on every rank instead of leaving the others blocked in the gather."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    err = None
    try:
        eng.upload(stream)
    except EngineError as ex:
        err = ex
    ready = [None] * world
    dist.all_gather_object(ready, (err is None, err is None and eng.ipc_ready()))
    if not all(r[0] for r in ready):
        raise err or EngineError("group-sharded run: another rank failed in engine_upload")
    if not all(r[1] for r in ready):
        handles = [None] * world
        dist.all_gather_object(handles, eng.ipc_export())
        if not eng.ipc_ready():
            eng.ipc_import(handles)
    torch.cuda.synchronize()
    dist.barrier()  # every rank's state is initialised before any kernel starts writing to its neighbours
    try:
        eng.run(stream)
        torch.cuda.synchronize()
    except EngineError as ex:
        err = ex
    dist.barrier()  # every rank's event loop is over: rank 0 may read the partitions
    try:
        if err is None:
            eng.finish(stream)  # rank 0: gathers every partition over the IPC mappings + final handlers
    except EngineError as ex:
        err = ex
    oks = [None] * world
    dist.all_gather_object(oks, err is None)  # also keeps the partitions alive until rank 0 has read them
    if not all(oks):
        raise err or EngineError("group-sharded run: another rank reported a device error")


def run_sharded_local(conf_path, world, lp_io_dir=None, **opts):
    """The group-sharded engine with all `world` ranks inside this process on ONE GPU (one engine and CUDA stream per rank,
    the grids sized to be co-resident): the cross-rank code path without a second GPU.  Returns the merged Summary."""
    import torch
    engines, streams = [], [torch.cuda.Stream() for _ in range(world)]
    try:
        for k in range(world):
            e = Engine()
            engines.append(e)
            e.setup(conf_path, make_opts(**opts))
            e.set_partition(k, world)
            e.set_ensemble(world)
            e.upload(streams[k].cuda_stream)
        arr = (_P * world)(*[e._h for e in engines])
        for e in engines:
            e._ck(e._lib.dfd_engine_attach_local_peers(e._h, arr, world))
        torch.cuda.synchronize()
        for k, e in enumerate(engines):
            e.run(streams[k].cuda_stream)
        torch.cuda.synchronize()
        for k in range(world - 1, -1, -1):
            engines[k].finish(streams[k].cuda_stream)
        if lp_io_dir:
            engines[0].lp_io_flush(lp_io_dir)
        return engines[0].summary()
    finally:
        for e in engines:
            e.close()


def run_sharded(conf_path, dist, lp_io_dir=None, stream=None, **opts):
    """Group-sharded run over the ranks of an initialised torch.distributed process group (one GPU per rank).
    Every rank simulates its own dragonfly groups; rank 0 returns the merged Summary (other ranks return None)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    eng = Engine()
    try:
        eng.setup(conf_path, make_opts(**opts))
        eng.set_partition(rank, world)
        sharded_tw_run(eng, dist, stream)
        out = None
        if rank == 0:
            if lp_io_dir:
                eng.lp_io_flush(lp_io_dir)
            out = eng.summary()
        dist.barrier()
        return out
    finally:
        eng.close()
