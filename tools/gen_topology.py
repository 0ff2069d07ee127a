#!/usr/bin/env python3
"""Generate dragonfly-dally link files and a matching .conf.

Produces the same binary formats the reference model reads
(`src/networks/model-net/dragonfly-dally.cxx:167-177`): the intra-group file is a stream of
`{int src, dest, type}` records for ONE group, the inter-group file a stream of
`{int src_gid, dest_gid}` records for the whole network.  The wiring rule is the one of the
reference generator (`scripts/dragonfly-dally/dragonfly-dally-topo-gen.py:19-31,120-156`):

  a = (radix+1)/2 routers per group, p = h = a/2, g = a*h/c + 1 groups (c = links between groups)
  intra: every ordered pair (s, d), s != d, in row-major order
  inter: group gi lists its a*h global ports; port l goes to the (l mod (g-1))-th group of the
         list "all other groups, rotated left by gi"; link i (global index over all groups) leaves
         router (i // h) mod a of its group and lands on router a-1-src_local of the peer group.

Usage: gen_topology.py <radix> <links_between_groups> <out_dir> [--name NAME] [--routing R] ...
"""
import argparse
import os
import struct


def dims(radix, conns):
    a = (radix + 1) // 2
    p = a // 2
    h = a // 2
    g = (h * a) // conns + 1
    return a, p, h, g


def intra_records(a, switch_conns=1):
    # switch_conns parallel local links between every ordered pair of routers (PARAMS num_parallel_switch_conns)
    return [(s, d, 0) for s in range(a) for d in range(a) if s != d for _ in range(switch_conns)]


def inter_records(a, h, g, conns):
    per_group = h * a
    recs = []
    i = 0
    for gi in range(g):
        others = [x for x in range(g) if x != gi]
        k = gi % len(others)
        others = others[k:] + others[:k]  # np.roll(others, -gi)
        for l in range(per_group):
            dest_g = others[l % len(others)]
            src_l = (i // h) % a
            dst_l = a - src_l - 1
            recs.append((src_l + gi * a, dst_l + dest_g * a))
            i += 1
    return recs


CONF_TEMPLATE = """LPGROUPS
{{
   MODELNET_GRP
   {{
      repetitions="{reps}";
      nw-lp="{p}";
      modelnet_dragonfly_dally="{p}";
      modelnet_dragonfly_dally_router="1";
   }}
}}
PARAMS
{{
   packet_size="{packet_size}";
   modelnet_order=( "dragonfly_dally","dragonfly_dally_router" );
   modelnet_scheduler="fcfs";
   chunk_size="{chunk_size}";
   num_routers="{a}";
   num_groups="{g}";
   local_vc_size="{local_vc}";
   global_vc_size="{global_vc}";
   cn_vc_size="{cn_vc}";
   local_bandwidth="{bw}";
   global_bandwidth="{bw}";
   cn_bandwidth="{bw}";
   message_size="736";
   num_cns_per_router="{p}";
   num_global_channels="{h}";
   intra-group-connections="{intra}";
   inter-group-connections="{inter}";
   routing="{routing}";
{extra}}}
"""


def write_config(out_dir, name, radix, conns, routing="minimal", packet_size=4096, chunk_size=4096,
                 local_vc=16384, global_vc=16384, cn_vc=32768, bw="2.0", extra=None, switch_conns=1):
    a, p, h, g = dims(radix, conns)
    os.makedirs(out_dir, exist_ok=True)
    intra = os.path.join(out_dir, f"{name}-intra")
    inter = os.path.join(out_dir, f"{name}-inter")
    with open(intra, "wb") as f:
        for r in intra_records(a, switch_conns):
            f.write(struct.pack("3i", *r))
    with open(inter, "wb") as f:
        for r in inter_records(a, h, g, conns):
            f.write(struct.pack("2i", *r))
    extra_lines = ""
    ex = dict(extra or {})
    if switch_conns != 1:
        ex["num_parallel_switch_conns"] = str(switch_conns)
    if routing == "prog-adaptive":
        ex.setdefault("adaptive_threshold", "16384")
        ex.setdefault("route_scoring_metric", "delta")
    for k, v in ex.items():
        extra_lines += f'   {k}="{v}";\n'
    conf = os.path.join(out_dir, f"{name}.conf")
    # an extra parameter that the template already sets replaces the template's value (a key never appears twice)
    overrides = {}
    for key, arg in (("local_vc_size", "local_vc"), ("global_vc_size", "global_vc"), ("cn_vc_size", "cn_vc")):
        if key in ex:
            overrides[arg] = ex.pop(key)
    local_vc, global_vc, cn_vc = (overrides.get("local_vc", local_vc), overrides.get("global_vc", global_vc), overrides.get("cn_vc", cn_vc))
    bws = {k: ex.pop(k) for k in ("local_bandwidth", "global_bandwidth", "cn_bandwidth") if k in ex}
    # a list value, e.g. router_buffer_snapshots=( "3000","8000" ), is written as it is
    extra_lines = "".join(f'   {k}={v};\n' if v.startswith("(") else f'   {k}="{v}";\n' for k, v in ex.items())
    with open(conf, "w") as f:
        # link files are written relative to the .conf (resolved against its directory)
        text = CONF_TEMPLATE.format(reps=a * g, p=p, a=a, g=g, h=h, packet_size=packet_size,
                                     chunk_size=chunk_size, local_vc=local_vc, global_vc=global_vc,
                                     cn_vc=cn_vc, bw=bw, intra=os.path.basename(intra),
                                     inter=os.path.basename(inter), routing=routing, extra=extra_lines)
        for k, v in bws.items():
            text = text.replace(f'{k}="{bw}";', f'{k}="{v}";')
        f.write(text)
    return conf, dict(a=a, p=p, h=h, g=g, routers=a * g, terminals=a * g * p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("radix", type=int)
    ap.add_argument("conns", type=int)
    ap.add_argument("out_dir")
    ap.add_argument("--name", default=None)
    ap.add_argument("--routing", default="minimal")
    ap.add_argument("--packet-size", type=int, default=4096)
    ap.add_argument("--chunk-size", type=int, default=4096)
    ap.add_argument("--switch-conns", type=int, default=1, help="parallel local links between each pair of routers of a group")
    ap.add_argument("--param", action="append", default=[], help="extra PARAMS key=value")
    args = ap.parse_args()
    a, p, h, g = dims(args.radix, args.conns)
    name = args.name or f"dfdally-{a * g * p}"
    extra = dict(kv.split("=", 1) for kv in args.param)
    conf, d = write_config(args.out_dir, name, args.radix, args.conns, routing=args.routing,
                           packet_size=args.packet_size, chunk_size=args.chunk_size, extra=extra, switch_conns=args.switch_conns)
    print(conf, d)


if __name__ == "__main__":
    main()
