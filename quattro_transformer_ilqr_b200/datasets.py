"""Training-set wire format: batched solver logs -> the reference's per-trajectory log entries.

The reference collects training data by running its MPC loop with ``enable_log=True`` and pickling
``iLQR_TF.logs`` - a list with one dict per iLQR iteration (quattro_ilqr_tf/quattro_ilqr_tf.py:453-466) - in
flushes, then re-pickling the entries one by one into a combined file
(examples/quadrotor/training/training_data_collection.py:202-218, :246-267) that ``transformer_training.py:9-42``
reads back into a DataFrame for ``TransformerILQR.fit`` / ``_create_dataset`` (transformer_ilqr.py:70-92).

A batched solve logs ONE dict per batch iteration with a leading trajectory axis (and an ``active`` mask).  The
helpers here cut those into the reference's single-trajectory entries, so a GPU sweep over thousands of initial
conditions produces a file the reference's training script consumes unchanged.
"""
from __future__ import annotations

import pickle

import numpy as np

# keys of a reference log entry whose values are per-time-step LISTS of arrays in the reference (B = 1 path)
_LIST_KEYS = ("u_seq", "new_u_seq", "k_seq", "K_seq", "k_seq_seg", "K_seq_seg")


def split_batched_logs(logs):
    """Batched logs (list of dicts with a leading B axis) -> list of single-trajectory entries in the reference's
    schema, ordered by trajectory then iteration; iterations in which a trajectory was already finished are dropped."""
    if not logs:
        return []
    B = len(np.asarray(logs[0]["current_cost"]).reshape(-1))
    out = []
    for b in range(B):
        for entry in logs:
            active = entry.get("active")
            if active is not None and not bool(np.asarray(active)[b]):
                continue
            found = bool(np.asarray(entry["found_update"])[b])
            e = {"iteration": int(entry["iteration"])}
            for key, val in entry.items():
                if key in ("iteration", "active"):
                    continue
                v = np.asarray(val)[b]
                if key == "found_update":
                    e[key] = found
                elif key in ("current_cost",):
                    e[key] = float(v)
                elif key in ("alpha", "new_cost"):
                    e[key] = float(v) if found else None
                elif key in ("new_x_seq",):
                    e[key] = v.copy() if found else None
                elif key in _LIST_KEYS:
                    e[key] = [v[t].copy() for t in range(v.shape[0])] if (found or key != "new_u_seq") else None
                else:
                    e[key] = v.copy()
            out.append(e)
    return out


def dump_entries(entries, path, append=False):
    """Write entries as a SEQUENCE of pickles, one per entry - the combined-file format of
    ``combine_logs_sequentially`` that ``load_ilqr_logs`` reads until EOF."""
    with open(path, "ab" if append else "wb") as f:
        for e in entries:
            pickle.dump(e, f)


def load_entries(path):
    """Inverse of :func:`dump_entries` (same loop as transformer_training.py:18-24)."""
    entries = []
    with open(path, "rb") as f:
        while True:
            try:
                entries.append(pickle.load(f))
            except EOFError:
                break
    return entries


def entries_to_dataframe(entries):
    import pandas as pd
    return pd.DataFrame([dict(e) for e in entries])
