"""BASELINE.json configs[2] and configs[4] at their named scale (bench tooling; the parity tests of the same configurations are
tests/test_configs_3_and_5.py).

  python tools/config_scale.py --config 3     variant-aware mode: 1 M PacBio-shape reads over a 24-contig, 3.1-Gb synthetic genome
                                              with a 5 M-record catalogue (85 % SNPs, 7.5 % insertions, 7.5 % deletions), one GPU
  torchrun ... tools/config_scale.py --config 5 [--reads 50000000]
                                              --dryRun inventory, then the full correction, of a PacBio + ONT mix on that genome,
                                              reads sharded over the ranks by shard_bounds, generated and corrected in slices so that
                                              host memory stays bounded

Every rank prints nothing; rank 0 prints one JSON line: set-up seconds (genome image, catalogue parse, table upload), reads/s end
to end through the C ABI with pinned host buffers (max over ranks), and the parity of a text sample against the oracle."""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422, 135086622,
        133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167, 46709983, 50818468,
        156040895, 57227415]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=(3, 5))
    ap.add_argument("--reads", type=int, default=0, help="total reads (default: 1 M for config 3, 50 M for config 5)")
    ap.add_argument("--slice", type=int, default=1000000, help="reads generated / corrected per step and rank")
    ap.add_argument("--variants", type=int, default=5000000)
    ap.add_argument("--scale", type=float, default=1.0, help="genome size factor (1.0 = 3.1 Gb); smaller for a dry run of this tool")
    ap.add_argument("--parity-reads", type=int, default=1500)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tools.workload import Workload, variants_from_reads
    from transcriptclean_b200 import engine as E, shard
    from transcriptclean_b200.genome import GenomeImage
    from transcriptclean_b200.tables import build_sj_tables, build_variant_tables
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    total = args.reads or (1000000 if args.config == 3 else 50000000)
    lens = [max(2000000, int(x * args.scale)) for x in HG38]
    names = ["chr%d" % (k + 1) for k in range(22)] + ["chrX", "chrY"]
    t0 = time.perf_counter()
    # ---- genome: 24 contigs; gene models (and therefore reads) on the first three, the rest is sequence only
    gi = GenomeImage(names, lens)
    works = []
    for c in range(24):
        w = Workload(seed=9000 + c, chrom_len=lens[c], n_genes=max(20, lens[c] // 83000) if c < 3 else 1, chrom=names[c])
        gi._put(int(gi.chrom_off[c]), w.genome)
        if c < 3:
            works.append(w)
        else:
            w.close(); del w
    t_genome = time.perf_counter() - t0
    d = tempfile.mkdtemp(prefix="tc_cfg_")
    sj_path = d + "/sj.tab"
    with open(sj_path, "w") as f:
        for w in works:
            for r in w.sj_rows():
                f.write("\t".join(str(x) for x in r) + "\n")
    # ---- catalogue (config 3): random records over all contigs + 30 % of the parity sample's own errors
    sample_w = works[0]
    scols, snm = sample_w.reads("pacbio" if args.config == 3 else "ont", args.parity_reads, seed_offset=777)
    vcf_path, t_vcf_write = None, 0.0
    if args.config == 3:
        t1 = time.perf_counter()
        vcf_path = d + "/catalogue.vcf"
        if rank == 0:
            rng = np.random.default_rng(5)
            n = args.variants
            contig = np.sort(rng.integers(0, 24, n))
            pos = (rng.random(n) * (np.asarray(lens)[contig] - 20)).astype(np.int64) + 5
            order = np.lexsort((pos, contig)); contig, pos = contig[order], pos[order]
            kind = rng.random(n)
            planted = variants_from_reads(sample_w, scols, 0, args.parity_reads, frac=0.3, seed=5)
            B = "ACGT"
            with open(vcf_path, "w") as f:      # (record order is free: processVCF keys a dict, tc_vcf_parse sorts)
                f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
                f.writelines("\t".join(str(x) for x in r) + "\t.\tPASS\t.\n" for r in planted)
                buf = []
                for k in range(n):
                    c, p, r = names[contig[k]], pos[k], B[k & 3]
                    if kind[k] < 0.85:
                        buf.append("%s\t%d\t.\t%s\t%s\t.\tPASS\t.\n" % (c, p, r, B[(k >> 2) & 3] + ("," + B[(k >> 4) & 3] if k % 167 == 0 else "")))
                    elif kind[k] < 0.925:
                        buf.append("%s\t%d\t.\t%s\t%s\t.\tPASS\t.\n" % (c, p, r, r + "ACGTA"[:1 + k % 5]))
                    else:
                        buf.append("%s\t%d\t.\t%s\t%s\t.\tPASS\t.\n" % (c, p, r + "TGCAT"[:1 + k % 5], r))
                    if len(buf) >= 200000:
                        f.writelines(buf); buf = []
                f.writelines(buf)
        t_vcf_write = time.perf_counter() - t1
        if world > 1:
            dist.barrier()
    # ---- engine + tables
    t2 = time.perf_counter()
    eng = E.Engine(local)
    eng.load_genome(gi)
    sj = build_sj_tables(sj_path, set(names[:3]), gi.index)
    eng.load_junctions(sj)
    t_sj = time.perf_counter() - t2
    t3 = time.perf_counter()
    n_rec = 0
    if vcf_path:
        var = build_variant_tables(vcf_path, 5, gi.index)          # all 24 contigs kept: the searches run over the full catalogue
        n_rec = var.n_records
        eng.load_variants(var)
    t_var = time.perf_counter() - t3
    # ---- parity sample: texts of the sample reads against the oracle (rank 0)
    parity = "not checked"
    if rank == 0 and args.parity_reads > 0:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import pipeline
        from transcriptclean_b200 import api
        lines = sample_w.sam_lines(scols, snm, 0, args.parity_reads)
        refs = api.Struct(); refs.genome = gi; refs.engine = eng
        opts = {"canonOnly": True} if args.config == 5 else {}
        got = api.correct_batch_text(lines, api.Struct(**opts), refs)
        want = pipeline.oracle_texts({names[0]: sample_w.genome.tobytes().decode()}, lines, sj_path, vcf_path, opts)
        diff = pipeline.first_diff(got, want)
        dry_ok = True
        if args.config == 5:
            gd = api.dry_run_text(lines, refs); wd = pipeline.oracle_texts({names[0]: sample_w.genome.tobytes().decode()}, lines, None, None, {}, dry=True)
            dry_ok = (gd["log"], gd["te"]) == (wd["log"], wd["te"])
        parity = ("ok: texts of %d reads equal the oracle's%s" % (args.parity_reads, " (dry run and correction)" if args.config == 5 else "")) if diff is None and dry_ok else "FAILED: %s" % diff
    # ---- the run: this rank's shard in slices
    lo, hi = shard.shard_bounds(total, world)[rank]
    mine = hi - lo
    stream = torch.cuda.ExternalStream(eng.stream())
    ms = {"dry": 0.0, "full": 0.0}
    done, vm, exact, te_rows = 0, 0, 0, 0
    step = 0
    while done < mine:
        n = min(args.slice, mine - done)
        w = works[step % 3]
        cidx = step % 3
        shape = "pacbio" if (args.config == 3 or step % 2 == 0) else "ont"
        cols, nm = w.reads(shape, n, seed_offset=1000 * rank + step)
        cols["chrom"] = np.where(cols["chrom"] == 0, cidx, cols["chrom"]).astype(np.int32)
        nt = max(1, (os.cpu_count() or 8) // world)                   # this rank's share of the host threads
        hcols = eng.pack_cigar(eng.pack_cols2(cols, n_threads=nt) or eng.pack_cols(cols, n_threads=nt), n_threads=nt)
        if "cig16" in hcols:
            hcols = {k: v for k, v in hcols.items() if k not in ("cig_op", "cig_len")}
        pinned = {}
        b = E._BatchIn(); b.n_reads = n
        for k, a in hcols.items():
            if not isinstance(a, np.ndarray):
                continue
            t = torch.empty(max(a.nbytes, 8), dtype=torch.uint8, pin_memory=True)
            if a.nbytes:
                t[:a.nbytes].copy_(torch.from_numpy(a.view(np.uint8).reshape(-1)))
            pinned[k] = t; setattr(b, k, t.data_ptr())
        b.seq_bits = int(hcols.get("seq_bits", 4)); b.n_exc = len(hcols["exc_pos"]) if "exc_pos" in hcols else 0
        b.n_cig_exc = len(hcols["cig_exc_idx"]) if "cig_exc_idx" in hcols else 0
        for mode in (("dry", "full") if args.config == 5 else ("full",)):
            if step < 2:
                (eng.raw_dryrun if mode == "dry" else eng.raw_correct)(b)      # warm-up of this mode's buffers (both read shapes)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()            # every rank times its call at the same moment: no rank's generator threads run beside it
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = (eng.raw_dryrun if mode == "dry" else eng.raw_correct)(b)
            e1.record(stream)
            torch.cuda.synchronize()
            ms[mode] += e0.elapsed_time(e1)
        cnt = np.frombuffer((C.c_char * (40 * n)).from_address(out.counters), dtype=np.int32).reshape(n, 10)
        vm += int(np.maximum(cnt[:, [2, 5]], 0).sum()); exact += int(eng.timing()["exact_nmmd_reads"])
        done += n; step += 1
        del pinned
    tm = torch.tensor([ms["dry"], ms["full"], float(mine), float(vm), float(exact)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tm.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tm.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = tm
    if rank == 0:
        line = {"config": args.config, "reads": total, "n_gpus": world, "genome_bases": int(sum(lens)), "contigs": 24,
                "catalogue_records": int(n_rec), "setup_s": {"genome_generate_and_pack": t_genome, "catalogue_file_write": t_vcf_write,
                                                              "sj_tables": t_sj, "catalogue_parse_and_upload": t_var},
                "e2e_reads_per_s": total / (float(mx[1]) * 1e-3), "e2e_ms": float(mx[1]),
                "variant_matched_indels": int(sm[3]), "exact_nmmd_reads": int(sm[4]), "parity": parity,
                "host_cores": os.cpu_count()}
        if args.config == 5:
            line["dry_run_reads_per_s"] = total / (float(mx[0]) * 1e-3); line["dry_run_ms"] = float(mx[0])
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
