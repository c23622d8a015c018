#!/usr/bin/env python3
"""Drop-in for the reference's src/plot_conservation.py (`memo view`) with the binning on the
B200 device path.

Same argv (src/plot_conservation.py:95-104).  preprocess_data -- the per-bin histogram of the
conservation vector, :46-65 -- runs as one kernel (memo_view_bins); the plot itself stays
plotnine, untouched (:67-90).

    python -m memo_b200.plot_conservation -i memo_conservation.txt -o out.png -n N -b BINS [-d DPI]
"""
import argparse


def parse_arguments(argv=None):
    ap = argparse.ArgumentParser(description="Plot k-mer conservation (binning on the B200 device path).")
    ap.add_argument("-i", "--in_file", dest="in_file", help="in file", required=True)
    ap.add_argument("-o", "--out_file", dest="out_file", help="output file", required=True)
    ap.add_argument("-n", "--ndocs", dest="n_docs", help="total number of genomes in the pangenome", required=True)
    ap.add_argument("-b", "--num_bins", dest="num_bins", help="number of histogram bins", required=True)
    ap.add_argument("-d", "--dpi", dest="dpi", default=600, help="Plot dpi", required=False)
    return ap.parse_args(argv)


def bin_composition(vals, n_docs, n_bins):
    """float64 [n_bins, n_docs + 1]: per position bin, the share of positions holding each
    conservation value (src/plot_conservation.py:49-58).  vals: device uint8 / int16 vector."""
    import numpy as np
    from . import api
    counts = api.view_bins(vals, n_docs, n_bins)
    sizes = np.diff(api.view_bin_edges(vals.numel(), n_bins))
    if (sizes == 0).any():
        raise ZeroDivisionError("division by zero")      # an empty bin: sum(Counter().values()) == 0 at :56
    return counts.astype(np.int64) / sizes[:, None]


def preprocess_data(path, n_docs, n_bins):
    """The data frame src/plot_conservation.py:46-65 hands to the plot: columns bin, No. Genomes,
    value; one row per (conservation value < n_docs, bin), value-major."""
    import numpy as np
    import pandas as pd
    import torch
    from . import io
    vals = io.read_int_text(path)                        # int() per stripped line (:49)
    if vals.size and vals.max() > 65535:
        raise ValueError("conservation values out of range")
    if not torch.cuda.is_available():
        from ._lib import MemoError
        raise MemoError("no CUDA device: memo_b200 has no CPU fallback")
    wide = vals.size and vals.max() > 255
    dev = torch.from_numpy(vals.astype(np.int16 if wide else np.uint8)).cuda()
    comp = bin_composition(dev, n_docs, n_bins)
    df = pd.DataFrame({
        "bin": np.tile(np.arange(n_bins, dtype=np.int64), n_docs),
        "No. Genomes": np.repeat(np.arange(n_docs, dtype=np.float64), n_bins),
        "value": comp[:, :n_docs].T.reshape(-1),
    })
    return df


def main(args):
    n_docs, n_bins, dpi = int(args.n_docs), int(args.num_bins), int(args.dpi)
    data = preprocess_data(args.in_file, n_docs, n_bins)
    # the plot is the reference's own (plotnine); only the binning moved to the device
    try:
        import plotnine  # noqa: F401
    except ImportError as e:
        raise ImportError("plotnine is needed for the plot itself (src/plot_conservation.py:67-90)") from e
    from plotnine import (aes, element_blank, element_line, element_text, geom_bar, ggplot, ggtitle,
                          scale_fill_gradient, scale_y_continuous, theme, themes, xlab, ylab)
    import numpy as np
    thm = themes.theme_bw(base_size=18, base_family="sans") + theme(
        legend_background=element_blank(), legend_key=element_blank(), panel_background=element_blank(),
        panel_border=element_blank(), strip_background=element_blank(), plot_background=element_blank(),
        panel_grid=element_blank(), axis_line=element_line(colour="black", size=1),
        axis_text_y=element_text(colour="black"), figure_size=[20, 4])
    p = (ggplot(data, aes(x="bin", y="value", fill="No. Genomes")) + geom_bar(stat="identity", width=1) +
         ggtitle("K-mer Conservation") + xlab("Genomic bin (n =" + str(n_bins) + ")") +
         ylab("Proportion of\nconserved k-mers") +
         scale_y_continuous(breaks=np.linspace(0, 1, 5), labels=["0", "0.25", "0.50", "0.75", "1"],
                            expand=(0, 0), limits=(0, 1)) +
         scale_fill_gradient(low="#000000", high="#c6dbef", limits=(1, n_docs - 1)) + thm)
    p.save(filename=args.out_file, dpi=dpi)


if __name__ == "__main__":
    main(parse_arguments())
