"""get_SJs_from_gtf (accessory_scripts/get_SJs_from_gtf.py:21-124 of the reference): splice junctions of a GTF
annotation in STAR SJ.out.tab format, the motif column looked up in the packed genome image (the same 2-bit
planes and literal runs the correction engine uses).  Host-side tool:

    python -m transcriptclean_b200.get_sjs_from_gtf --f annotation.gtf --g genome.fa --o SJs.tab [-m 21]

Same options, same rows, same order as the reference script: exons must be in transcript order, a junction is
written the first time its row is seen, introns shorter than --minIntronSize are dropped."""
from optparse import OptionParser

from .genome import GenomeImage

_MOTIF = {"GTAG": "21", "CTAC": "22", "GCAG": "23", "CTGC": "24", "ATAC": "25", "GTAT": "26"}


def get_options(argv=None):
    parser = OptionParser()
    parser.add_option("--f", dest="infile", help="Input GTF file", metavar="FILE", type="string", default="")
    parser.add_option("--g", dest="genomeFile", help="Reference genome", metavar="FILE", type="string", default="")
    parser.add_option("--minIntronSize", "-m", dest="minIntron", default=21,
                      help="Minimum size of intron to consider a junction. Default: 21 bp.")
    parser.add_option("--o", dest="outfile", help="output file", metavar="FILE", type="string", default="out.txt")
    return parser.parse_args(argv)[0]


def intron_motif(genome, chrom, start, end):
    """getIntronMotif (get_SJs_from_gtf.py:44-66): first two + last two intron bases, upper-cased."""
    motif = (genome.get_seq(chrom, start, start + 1) + genome.get_seq(chrom, end - 1, end)).upper()
    return _MOTIF.get(motif, "20")


def format_sj(exon, prev_exon_end, genome, min_intron):
    """formatSJOutput (get_SJs_from_gtf.py:21-42).  `exon` = the tab-split GTF line of the current exon."""
    chrom, strand = exon[0], exon[6]
    if strand == "+":
        code, a, b = "1", int(prev_exon_end) + 1, int(exon[3]) - 1
    elif strand == "-":
        code, a, b = "2", int(exon[4]) + 1, int(prev_exon_end) - 1
    else:
        raise ValueError("exon without a strand: " + "\t".join(exon[:9]))    # the reference dies here (unbound names)
    intron_motif(genome, chrom, a, b)          # the reference looks the motif up before the size test (fetch errors first)
    if abs(b - a + 1) < min_intron:
        return None
    return "\t".join([chrom, str(a), str(b), code, intron_motif(genome, chrom, a, b), "1", "NA", "NA", "NA"])


def junction_rows_device(gtf_lines, genome, engine, min_intron=21):
    """junction_rows with the motif column from the device: the candidate introns are collected first, their motifs come from
    one call of the junction kernel's motif fetch (Engine.intron_motifs; the engine holds `genome`), then the rows are written
    with the reference's rules (first occurrence wins, short introns dropped after the lookup)."""
    cand, prev_tid, prev_end = [], "", 0
    for line in gtf_lines:
        line = line.strip()
        if line.startswith("#"):
            continue
        info = line.split("\t")
        if info[2] != "exon" or "transcript_id" not in info[-1]:
            continue
        tid = info[-1].split("transcript_id ")[1].split('"')[1]
        strand = info[6]
        if tid != prev_tid:
            prev_tid = tid
        else:
            if strand == "+":
                cand.append((info[0], "1", int(prev_end) + 1, int(info[3]) - 1))
            elif strand == "-":
                cand.append((info[0], "2", int(info[4]) + 1, int(prev_end) - 1))
            else:
                raise ValueError("exon without a strand: " + "\t".join(info[:9]))
        prev_end = info[4] if strand == "+" else info[3]
    for c in cand:
        if c[0] not in genome.index:
            raise KeyError("Requested rname %s does not exist!" % c[0])
    codes = engine.intron_motifs([genome.index[c[0]] for c in cand], [c[2] for c in cand], [c[3] for c in cand])
    seen = set()
    for (chrom, code, a, b), mc in zip(cand, codes.tolist()):
        if mc < 0:
            raise IndexError("Requested start coordinate must be greater than 0.")
        if abs(b - a + 1) < min_intron:
            continue
        row = "\t".join([chrom, str(a), str(b), code, str(20 + mc), "1", "NA", "NA", "NA"])
        if row not in seen:
            seen.add(row)
            yield row


def junction_rows(gtf_lines, genome, min_intron=21):
    """main() of the reference script (get_SJs_from_gtf.py:68-120) as a generator of output rows."""
    seen = set()
    prev_tid, prev_end = "", 0
    for line in gtf_lines:
        line = line.strip()
        if line.startswith("#"):
            continue
        info = line.split("\t")
        if info[2] != "exon":
            continue
        desc = info[-1]
        if "transcript_id" not in desc:
            continue
        tid = desc.split("transcript_id ")[1].split('"')[1]
        strand = info[6]
        if tid != prev_tid:
            prev_tid = tid
        else:
            row = format_sj(info, prev_end, genome, min_intron)
            if row is not None and row not in seen:
                seen.add(row)
                yield row
        prev_end = info[4] if strand == "+" else info[3]


def main(argv=None):
    o = get_options(argv)
    genome = GenomeImage.from_fasta_cached(o.genomeFile)
    engine = None
    try:                                   # the motif fetch of the junction kernel when a GPU is there, the host image otherwise
        from .engine import Engine
        engine = Engine(0)
        engine.load_genome(genome)
    except Exception:
        engine = None
    with open(o.infile) as f, open(o.outfile, "w") as out:
        rows = junction_rows_device(f, genome, engine, int(o.minIntron)) if engine is not None else junction_rows(f, genome, int(o.minIntron))
        for row in rows:
            out.write(row + "\n")
    if engine is not None:
        engine.close()


if __name__ == "__main__":
    main()
