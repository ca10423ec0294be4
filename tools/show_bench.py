import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6),
              {k.replace("_ms", ""): round(v, 2) for k, v in d["kernel_ms_per_step"].items()},
              "frac %.3f" % d["roofline"]["frac"], "exact", d["exact_nmmd_reads_per_step"], "ascii", d.get("ascii_path_reads_per_step"))
    except Exception as e:
        print(f, "ERR", e)
