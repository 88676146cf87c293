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
