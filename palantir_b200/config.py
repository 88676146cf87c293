"""Module-level switches honoured by the hot path (mirror of src/palantir/config.py:16-39;
the matplotlib rcParams the reference sets on import are plotting-only and not mirrored)."""

# Store DataFrames (True) or bare arrays + a ``<key>_columns`` entry (False) in AnnData.obsm.
SAVE_AS_DF = True

# Convert integer cell identifiers to str when obs_names are strings.
AUTO_CONVERT_CELL_IDS_TO_STR = True

# Warn when such a conversion happens or when requested cells are missing.
WARN_ON_CELL_ID_CONVERSION = True

# Accepted for signature compatibility; the B200 path runs no joblib workers.
JOBLIB_BACKEND = None

# "scanpy" | "sklearn".  Both names run the exact-kNN ("sklearn") semantics here: the
# reference's scanpy branch is an approximate pynndescent search and is not reproducible.
KERNEL_BACKEND = "scanpy"

# Multi-GPU runs (one process per GPU): the two sparse results of run_diffusion_maps (kernel and T, 1.3 GB at
# 1 M cells) are brought to the host on every rank, as a single-process caller expects.  False: only rank 0
# materialises them ("kernel" / "T" are None elsewhere; everything else is returned on every rank) -- eight
# processes pulling them through one host at the same time cost more than the whole device step.
HOST_SPARSE_ON_ALL_RANKS = True

# Reuse the device originals of arrays this package returned (eigenvectors, multiscale matrix, kernel) when the
# very same host objects come back in the next call, instead of uploading them again.  Off by default: an in-place
# edit of a few entries of such an array cannot be detected without re-reading all of it (see _devcache.py).
REUSE_DEVICE_RESULTS = False
