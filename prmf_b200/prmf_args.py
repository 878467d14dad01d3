"""Command-line schema of prmf_runner.py: the reference's 19 flags (prmf/prmf_args.py:7-26), verbatim in
name, default and meaning, plus `--normalize` (accepted as a no-op: the reference README advertises it
but argparse rejects it; normalisation is on unless `--no-normalize` is given)."""


def _bool(text):
    """The reference declares `--high-dimensional` with type=bool, so every non-empty string -- including
    "False" -- parses as True (prmf/prmf_args.py:21).  Here the usual spellings of false work."""
    if isinstance(text, bool):
        return text
    return str(text).strip().lower() not in ("", "0", "false", "no", "off")


def add_prmf_arguments(parser):
    add = parser.add_argument
    add("--data", type=str, required=True, help="delimited text file holding the samples x genes matrix")
    add("--manifolds", nargs='+', help="pathway graphs, one .graphml file each; their node names must occur in the nodelist")
    add("--manifolds-file", help="text file listing the .graphml files, one path per line (instead of --manifolds)")
    add("--manifolds-init", nargs='*',
        help="seed the factors from pathways: fewer files than -k (or none) are topped up with randomly chosen ones, "
             "exactly -k files are used as given, more than -k are subsampled")
    add("--node-attribute", default=None, help="graphml node attribute that carries the gene name (parsed, as in the reference)")
    add("--outdir", type=str, required=True, help="where U.csv, V.csv and obj.txt are written")
    add("--nodelist", type=str, help="gene order of the matrix columns, whitespace separated; default: the header of --data")
    add("--k-latent", "-k", default=6, type=int, help="number of factors")
    add("--tolerence", type=float, default=1e-3)
    add("--seed", default=None)
    add("--gamma", default=1.0, type=float, help="weight of the Laplacian (manifold) term before rescaling by ||X||/k; default 1.0")
    add("--delta", default=1.0, type=float, help="weight of the penalty for ignoring the manifold before rescaling; default 1.0")
    add("--tradeoff", default=-1, type=float,
        help="in [0,1]: re-derive gamma = delta from the last objective after every inner step (larger favours the "
             "manifold term); -1 (default) keeps them fixed")
    add("--high-dimensional", default=None, type=_bool,
        help="true: transpose the input if it has more rows than columns; false: if it has fewer (the reference's rule, "
             "prmf_runner.py:946-952).  Default: the matrix is taken as written, samples x genes -- the reference's default "
             "of true would factor the transpose of a recount2-shape file and then fail writing U.csv (SURVEY 0.7)")
    add("--no-normalize", action='store_true', help="skip the quantile normalisation of the data")
    add("--normalize", action='store_true', help="no-op kept for the reference README's spelling; normalisation is on by default")
    add("--delimiter", default=",", help="field separator of --data")
    add("--m-samples", type=int, help="parsed and ignored, as in the reference (its value never reaches read_csv, prmf_runner.py:933-942)")
    add("--cross-validation", "-c", type=float, help="hold out this fraction of the samples and report their reconstruction error")
    add("--verbose", "-v", action='store_true', help="print the pathway assignments and objective parts of every iteration")
