"""Cell-identifier reconciliation used by run_palantir (host side only).

Mirror of ``normalize_cell_identifiers`` (src/palantir/validation.py:162-308):
accepts a list / dict / Series / Index / array of cell ids, reconciles int-vs-str
mismatches against ``obs_names`` according to ``config``, warns the way the
reference does and returns ``(id -> annotation dict, n_requested, n_found)``.
"""
from __future__ import annotations

import warnings
from typing import Dict, Tuple

import numpy as np
import pandas as pd

from . import config


def _is_int(v) -> bool:
    return isinstance(v, (int, np.integer))


def normalize_cell_identifiers(cells, obs_names: pd.Index, context: str = "cell selection") -> Tuple[Dict, int, int]:
    if isinstance(cells, dict):
        table = dict(cells)
    elif isinstance(cells, pd.Series):
        table = dict(cells)
    elif isinstance(cells, (list, pd.Index, np.ndarray)):
        table = {c: "" for c in cells}
    else:
        raise ValueError(
            f"Invalid cells format for {context}. Expected list, dict, pd.Series, "
            f"pd.Index, or np.ndarray, got {type(cells).__name__}")
    if not table:
        raise ValueError(f"No cells provided for {context}")

    n_requested = len(table)
    hits_before = sum(1 for c in table if c in obs_names)
    first_cell, first_obs = next(iter(table)), obs_names[0]
    cell_type, obs_type = type(first_cell).__name__, type(first_obs).__name__

    mismatch = None
    if _is_int(first_cell) and isinstance(first_obs, str):
        mismatch = "Cell identifiers are integers but data.obs_names contains strings"
    elif isinstance(first_cell, str) and _is_int(first_obs):
        mismatch = "Cell identifiers are strings but data.obs_names contains integers"

    converted, how = False, ""
    if mismatch and config.AUTO_CONVERT_CELL_IDS_TO_STR:
        try:
            if _is_int(first_cell):
                table = {str(k): v for k, v in table.items()}
                converted, how = True, "Converting cell identifiers from int to str"
            elif _is_int(first_obs):
                as_str = obs_names.astype(str)
                if sum(1 for c in table if c in as_str) > hits_before:
                    obs_names = as_str
                    converted, how = True, "Treating data.obs_names as strings for matching"
        except (ValueError, TypeError):
            pass

    n_found = sum(1 for c in table if c in obs_names)

    if config.WARN_ON_CELL_ID_CONVERSION:
        if mismatch:
            if converted and n_found > 0:
                warnings.warn(
                    f"{context}: {mismatch}. {how} (config.AUTO_CONVERT_CELL_IDS_TO_STR=True). "
                    f"Found {n_found}/{n_requested} cells after conversion. "
                    f"To disable this warning, set config.WARN_ON_CELL_ID_CONVERSION=False. "
                    f"To fix this issue, ensure your cell identifiers match the type of data.obs_names: "
                    f"use pd.Series(values, index=data.obs_names[your_indices]) or ensure consistent types.",
                    UserWarning, stacklevel=3)
            elif not converted:
                warnings.warn(
                    f"{context}: {mismatch}. "
                    f"Automatic conversion is disabled (config.AUTO_CONVERT_CELL_IDS_TO_STR=False). "
                    f"Only {n_found}/{n_requested} cells matched. "
                    f"To enable automatic conversion, set config.AUTO_CONVERT_CELL_IDS_TO_STR=True. "
                    f"To fix this issue manually, ensure your cell identifiers match the type of data.obs_names.",
                    UserWarning, stacklevel=3)
        if n_found == 0:
            warnings.warn(
                f"{context}: None of the {n_requested} requested cells were found in data.obs_names. "
                f"Your cells have type {cell_type}, data.obs_names has type {obs_type}. "
                f"When using pd.Series, the index should contain actual cell identifiers from data.obs_names, "
                f"not positional indices. Example: pd.Series(['label1', 'label2'], index=['cell_A', 'cell_B'])",
                UserWarning, stacklevel=3)
        elif n_found < n_requested:
            warnings.warn(
                f"{context}: Only {n_found}/{n_requested} requested cells were found in data.obs_names "
                f"({n_requested - n_found} cells missing). Check that cell identifiers match data.obs_names exactly.",
                UserWarning, stacklevel=3)
    return table, n_requested, n_found
