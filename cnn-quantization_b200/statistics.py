"""Offline calibration statistics (`-sm collect` / `-sm use`): mirrors of the reference's
``StatisticManager`` (per tensor, CSV; statistic_manager.py:15-178) and ``StatisticManagerPerChannel`` (per channel,
pickle; statistic_manager_perchannel.py:17-174) with the same on-disk formats, so statistics collected by either
implementation can be used by the other:

    <base>/statistics/<folder>/<id>.csv                         one row per batch, columns = statistic names
    <base>/statistics/<folder>/<folder>_summary.csv            index = id, columns internal_name, {min,mean,max}_<stat>, dim
    <base>/statistics/per_channel/<folder>/<folder>_statistics_perchannel_summary.pkl
                                                               {id: DataFrame[{min,mean,max}_<stat>], one row per channel}

``base`` defaults to ``~/mxt-sim`` like the reference (override with the constructor argument or $FQB200_STATS_DIR).

Collect mode is an offline, one-time pass.  The five statistics the quantizers consume in use mode (min, max, mean,
b, std) come from ONE statistics-only launch of the fused kernel per hooked tensor; the diagnostic columns the
reference also writes (kurtosis, mean_abs, std_pos) are computed with plain torch ops; the error columns
(mse_*/cos_*) are NaN exactly as in the reference's own collect runs (it never passes quantized tensors there).
"""
import os
import pickle
import re
import shutil

import numpy as np
import torch

from . import ops

__all__ = ["StatisticManager", "StatisticManagerPerChannel", "default_base_dir"]


def default_base_dir():
    return os.environ.get("FQB200_STATS_DIR") or os.path.join(os.path.expanduser("~"), "mxt-sim")


def sorted_nicely(keys):
    """Natural sort (conv2 before conv10), utils/misc.py:77-88."""
    conv = lambda t: int(t) if t.isdigit() else t
    return sorted(keys, key=lambda k: [conv(c) for c in re.split("([0-9]+)", k)])


_ERR_COLUMNS = ["mse_lowp", "mse_gaus", "mse_laplace", "cos_lowp", "cos_gaus", "cos_laplace"]


class StatisticManager(object):
    """Per-tensor statistics (statistic_manager.py:15-178)."""

    def __init__(self, folder, load_stats, stats=None, batch_avg=False, kld_threshold=False, collect_err=True, base_dir=None):
        if kld_threshold:
            raise NotImplementedError("KLD thresholds are outside the hot-path scope (SURVEY.md section 2, #9)")
        self.name = folder
        self.folder = os.path.join(base_dir or default_base_dir(), "statistics", folder)
        self.stats_names = list(stats) if stats is not None else ["max", "min", "std", "mean", "kurtosis", "mean_abs", "b", "dim"]
        self.batch_avg = batch_avg
        if collect_err:
            self.stats_names += _ERR_COLUMNS
        self.stats = {}
        self.metadata = {}
        self.save_stats = not load_stats
        self.stats_df = None
        if load_stats:
            import pandas as pd
            path = os.path.join(self.folder, "%s_summary.csv" % self.name)
            if not os.path.exists(path):
                raise FileNotFoundError("no collected statistics at %s (run with stats_mode='collect' first)" % path)
            self.stats_df = pd.read_csv(path, index_col=0)

    # -- collect ---------------------------------------------------------------------------------------
    def save_tensor_stats(self, tensor, tag, id, tensors_q=None, force_global_min_max=False):
        """One row of statistics for this batch (statistic_manager.py:47-122)."""
        t = tensor.detach().contiguous()
        n = t.shape[0]
        st = ops.fused(t, (1, 1, t.numel()), stats_only=True)[0]  # min max mean b std over the whole tensor
        glob = {"min": st[0], "max": st[1], "mean": st[2], "b": st[3], "std": st[4]}
        if self.batch_avg and not force_global_min_max:
            per = ops.fused(t, (1, n, t.numel() // n), stats_only=True)
            glob["min"], glob["max"] = per[:, 0].mean(), per[:, 1].mean()
        row = []
        flat = t.view(-1)
        for sn in self.stats_names:
            if sn in glob:
                v = glob[sn]
            elif sn == "kurtosis":
                v = torch.mean(((flat - glob["mean"]) / glob["std"]) ** 4) - 3
            elif sn == "mean_abs":
                v = torch.mean(flat.abs())
            elif sn == "dim":
                v = flat.numel()
            else:  # mse_* / cos_*: the reference writes NaN when no quantized tensors are handed in
                v = float("nan")
            row.append(float(v))
        arr = np.asarray(row, dtype=np.float64).reshape(1, -1)
        if id in self.stats:
            self.stats[id] = np.concatenate([self.stats[id], arr])
        else:
            self.stats[id] = arr
            self.metadata[id] = tag

    # -- use -------------------------------------------------------------------------------------------
    def get_tensor_stat(self, id, stat, kind="mean"):
        if self.stats_df is None:
            return None
        return self.stats_df.loc[id, "%s_%s" % (kind, stat)]

    def get_tensor_stats(self, id, kind=None):
        kind = kind or {"min": "mean", "max": "mean", "mean": "mean", "std": "mean", "mean_abs": "mean", "b": "mean"}
        if self.stats_df is None:
            return (None,) * 6
        return tuple(self.stats_df.loc[id, "%s_%s" % (kind[s], s)] for s in ("min", "max", "mean", "std", "mean_abs", "b"))

    # -- persistence (statistic_manager.py:146-178) ----------------------------------------------------------
    def __exit__(self, *args):
        if not self.save_stats:
            return
        import pandas as pd
        if os.path.exists(self.folder):
            shutil.rmtree(self.folder)
        os.makedirs(self.folder)
        frames = {}
        for s_id, data in self.stats.items():
            df = pd.DataFrame(columns=self.stats_names, data=data)
            df.to_csv(os.path.join(self.folder, "%s.csv" % s_id), index=False)
            frames[s_id] = df
        cols = []
        for c in self.stats_names:
            cols += ["min_%s" % c, "mean_%s" % c, "max_%s" % c]
        summary = pd.DataFrame(columns=["internal_name"] + cols)
        for s_id in sorted_nicely(frames.keys()):
            summary.loc[s_id, "internal_name"] = self.metadata[s_id]
            for c in self.stats_names:
                summary.loc[s_id, "min_%s" % c] = frames[s_id][c].min()
                summary.loc[s_id, "mean_%s" % c] = frames[s_id][c].mean()
                summary.loc[s_id, "max_%s" % c] = frames[s_id][c].max()
            summary.loc[s_id, "dim"] = frames[s_id]["dim"][0] if "dim" in frames[s_id] else np.nan
        summary.to_csv(os.path.join(self.folder, "%s_summary.csv" % self.name), index=True)

    def __enter__(self):
        return self


class StatisticManagerPerChannel(object):
    """Per-channel statistics of [N, C, H, W] activations (statistic_manager_perchannel.py:17-174)."""

    def __init__(self, folder, load_stats, stats=None, batch_avg=False, collect_err=False, base_dir=None):
        self.name = folder
        self.folder = os.path.join(base_dir or default_base_dir(), "statistics/per_channel", folder)
        self.stats_names = list(stats) if stats is not None else ["max", "min", "std", "mean", "kurtosis", "b", "std_pos"]
        self.batch_avg = batch_avg
        if collect_err:
            self.stats_names += _ERR_COLUMNS
        self.save_stats = not load_stats
        self.stats = {}
        if load_stats:
            path = os.path.join(self.folder, "%s_statistics_perchannel_summary.pkl" % self.name)
            if not os.path.exists(path):
                raise FileNotFoundError("no collected per-channel statistics at %s" % path)
            with open(path, "rb") as f:
                self.stats = pickle.load(f)

    def save_tensor_stats(self, tensor, tag, id, tensors_q=None, force_global_min_max=False):
        """Per-channel rows for this batch (statistic_manager_perchannel.py:45-122); FC / 1x1 tensors are skipped."""
        if tensor.dim() < 3 or (tensor.shape[2] == 1 and tensor.shape[3] == 1):
            return
        t = tensor.detach().contiguous()
        n, c = t.shape[0], t.shape[1]
        hw = t.numel() // (n * c)
        st = ops.fused(t, (n, c, hw), stats_only=True)  # [C, 12]: min max mean b std ...
        vals = {"min": st[:, 0], "max": st[:, 1], "mean": st[:, 2], "b": st[:, 3], "std": st[:, 4]}
        if not force_global_min_max:
            per = ops.fused(t, (1, n * c, hw), stats_only=True).view(n, c, -1)  # per (n, c)
            if self.batch_avg:
                vals["max"], vals["min"] = per[:, :, 1].mean(0), per[:, :, 0].mean(0)
            else:
                vals["min"] = per[:, :, 0].min(0)[0]  # the reference's non-averaged min is the min over n of per-(n,c) minima
        tc = None
        for sn in self.stats_names:
            if sn in vals:
                v = vals[sn]
            elif sn in ("kurtosis", "std_pos"):
                if tc is None:
                    tc = t.transpose(0, 1).reshape(c, -1)  # offline diagnostics only: a transposed copy is fine here
                if sn == "kurtosis":
                    v = torch.mean(((tc - vals["mean"].unsqueeze(-1)) / vals["std"].unsqueeze(-1)) ** 4, dim=-1) - 3
                else:
                    v = torch.std(torch.relu(tc), dim=-1, unbiased=True)
            else:
                continue  # mse_* / cos_* need quantized tensors: skipped like the reference does
            v = v.detach().cpu().numpy()
            entry = self.stats.setdefault(id, {})
            entry[sn] = v if sn not in entry else np.vstack([entry[sn], v])

    def get_tensor_stat(self, id, stat, kind="mean"):
        if self.stats is None:
            return None
        return self.stats[id]["%s_%s" % (kind, stat)]

    def __exit__(self, *args):
        if not self.save_stats:
            return
        import pandas as pd
        if os.path.exists(self.folder):
            shutil.rmtree(self.folder)
        os.makedirs(self.folder)
        cols = []
        for c in self.stats_names:
            cols += ["min_%s" % c, "mean_%s" % c, "max_%s" % c]
        summary = {}
        for layer, entry in self.stats.items():
            df = pd.DataFrame(columns=cols)
            for s in self.stats_names:
                if s in entry:
                    t = entry[s]
                    multi = len(t.shape) > 1
                    df["min_%s" % s] = t.min(axis=0) if multi else [t.min(axis=0)]
                    df["mean_%s" % s] = t.mean(axis=0) if multi else [t.mean(axis=0)]
                    df["max_%s" % s] = t.max(axis=0) if multi else [t.max(axis=0)]
            summary[layer] = df
        with open(os.path.join(self.folder, "%s_statistics_perchannel_summary.pkl" % self.name), "wb") as f:
            pickle.dump(summary, f)

    def __enter__(self):
        return self
