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
    parser.add_argument("--data", type=str, required=True, help="n_obs x n_features matrix")
    parser.add_argument("--manifolds", nargs='+', help="graphml files to use as manifold. Node identifiers must appear in nodelist.")
    parser.add_argument("--manifolds-file", help="A file containing newline-delimited filepaths which are used as graphml files as in <manifolds>")
    parser.add_argument("--manifolds-init", nargs='*', help="If provided, use this list of manifolds to initialize PRMF (see the reference for the three cases).")
    parser.add_argument("--node-attribute", help="Relabel nodes in manifolds/graphs so that their node identifiers come from this node attribute.", default=None)
    parser.add_argument("--outdir", type=str, required=True, help="Directory containing results")
    parser.add_argument("--nodelist", type=str, help="Association of node identifier to matrix indexes. If not provided, inferred from the header in <--data>.")
    parser.add_argument("--k-latent", "-k", default=6, help="Number of latent factors", type=int)
    parser.add_argument("--tolerence", type=float, default=1e-3)
    parser.add_argument("--seed", default=None)
    parser.add_argument("--gamma", default=1.0, help="Tradeoff between reconstruction error and manifold regularization term; Default = 1.0", type=float)
    parser.add_argument("--delta", default=1.0, help="Regularization parameter for penalty for ignoring manifold; Default = 1.0", type=float)
    parser.add_argument("--tradeoff", default=-1, type=float, help="If set, automatically update gamma and delta from the previous iteration's objective parts. Must be in [0,1]; -1 disables. Default = -1.")
    parser.add_argument("--high-dimensional", default=True, type=_bool, help="If True, ensure that <data> is of shape m x n with m < n ; otherwise ensure m > n. Default = True.")
    parser.add_argument("--no-normalize", action='store_true', help="If flag is provided, don't quantile normalize the data")
    parser.add_argument("--normalize", action='store_true', help="Accepted for compatibility with the reference README; normalisation is the default.")
    parser.add_argument("--delimiter", default=",", help="Field delimiter in <--data>")
    parser.add_argument("--m-samples", help="If provided, only use the first <--m-samples> rows in <--data>", type=int)
    parser.add_argument("--cross-validation", "-c", type=float, help="Fraction of the samples to hold out and measure model performance with")
    parser.add_argument("--verbose", "-v", action='store_true', help="Report more information during each iteration")
