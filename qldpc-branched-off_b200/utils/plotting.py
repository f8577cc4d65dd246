"""Result plots with the reference's function names (``src/utils/plotting.py:5,37,92``) so that an unmodified
``main.py`` runs against this package.  Presentation only: when matplotlib is not installed the figures are
skipped, the numeric return value of ``plot_alpha_linearity`` (R^2 of a straight-line fit per alpha sequence,
stored by main.py) is still computed."""
import math

import numpy as np

_COLORS = ["#2E72AE", "#64B791", "#DBA142", "#000000", "#E17792"]


def _plt():
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        return None


def _line_r2(y):
    y = np.asarray(y, dtype=float)
    if y.size < 2:
        return float("nan"), None
    x = np.arange(1, y.size + 1, dtype=float)
    fit = np.polyval(np.polyfit(x, y, 1), x)
    ss_res, ss_tot = np.sum((y - fit) ** 2), np.sum((y - y.mean()) ** 2)
    return 1.0 - (ss_res / ss_tot if ss_tot > 0 else np.nan), fit


def plot_simulation_results(results, filename="simulation_results.png"):
    plt = _plt()
    if plt is None:
        print("matplotlib not available: skipping", filename)
        return
    plt.figure(figsize=(10, 7))
    for i, (name, data) in enumerate(results.items()):
        ps = np.array(sorted(data.keys()), dtype=float)
        lers = np.array([data[p]["logical_error_rate"] for p in sorted(data.keys())], dtype=float)
        color = _COLORS[i % len(_COLORS)]
        plt.loglog(ps, lers, "o", label=f"n={name}", color=color)
        ok = (ps > 0) & (lers > 0)
        if ok.sum() >= 2:
            slope, icpt = np.polyfit(np.log10(ps[ok]), np.log10(lers[ok]), 1)
            xs = np.linspace(-4, np.log10(ps.max()), 200)
            plt.loglog(10 ** xs, 10 ** (slope * xs + icpt), "-", color=color)
    plt.xlabel("Physical Error Rate p"); plt.ylabel("Logical Error Rate LER")
    plt.xlim(1e-4, 1e-2); plt.grid(True, which="both", ls="-", alpha=0.5); plt.legend()
    plt.title("Spatio-Temporal Decoding Performance")
    plt.savefig(filename, dpi=300); plt.close()
    print(f"\nResults saved to {filename}")


def _codes_with_alpha(results):
    return [name for name, data in results.items() if any("alpha_values_z" in r for r in data.values())]


def plot_alpha_comparison(results, filename="alpha_comparison.png"):
    codes = _codes_with_alpha(results)
    if not codes:
        print("No autoregressive alpha values found to plot.")
        return
    plt = _plt()
    if plt is None:
        print("matplotlib not available: skipping", filename)
        return
    ncols = 2 if len(codes) > 1 else 1
    nrows = math.ceil(len(codes) / ncols)
    fig, axes = plt.subplots(nrows, ncols, figsize=(7 * ncols, 4 * nrows), squeeze=False)
    for ax, name in zip(axes.flat, codes):
        first = True
        for p in sorted(results[name].keys()):
            res = results[name][p]
            if "alpha_values_z" not in res:
                continue
            az = np.asarray(res["alpha_values_z"], dtype=float)
            it = np.arange(1, az.size + 1)
            if first:
                ax.plot(it, 1.0 - 2.0 ** (-it.astype(float)), "k--", label="dynamical")
                first = False
            ax.plot(it, az, label=f"p={p} Z")
            if "alpha_values_x" in res:
                ax.plot(it, np.asarray(res["alpha_values_x"], dtype=float), ":", label=f"p={p} X")
        ax.set_title(f"n={name}"); ax.set_xlabel("Iteration"); ax.set_ylabel("Alpha"); ax.grid(True, ls="-", alpha=0.4); ax.legend(fontsize=8)
    for idx in range(len(codes), nrows * ncols):
        fig.delaxes(axes.flat[idx])
    plt.tight_layout(); plt.savefig(filename, dpi=300); plt.close()
    print(f"\nAlpha comparison plot saved to {filename}")


def plot_alpha_linearity(results, filename="alpha_linearity.png"):
    """Returns r2_values[code][p] = {"z": r2_z, "x": r2_x} like the reference (plotting.py:92-161)."""
    r2_values = {}
    codes = _codes_with_alpha(results)
    if not codes:
        print("No autoregressive alpha values found to plot.")
        return r2_values
    plt = _plt()
    axes = None
    if plt is not None:
        ncols = 2 if len(codes) > 1 else 1
        nrows = math.ceil(len(codes) / ncols)
        fig, axes = plt.subplots(nrows, ncols, figsize=(7 * ncols, 4 * nrows), squeeze=False)
    for k, name in enumerate(codes):
        ax = axes.flat[k] if axes is not None else None
        r2_values.setdefault(name, {})
        for p in sorted(results[name].keys()):
            res = results[name][p]
            if "alpha_values_z" not in res:
                continue
            r2z, fz = _line_r2(res["alpha_values_z"])
            r2x, fx = (float("nan"), None)
            if "alpha_values_x" in res:
                r2x, fx = _line_r2(res["alpha_values_x"])
            r2_values[name][p] = {"z": r2z, "x": r2x}
            if ax is not None:
                it = np.arange(1, len(res["alpha_values_z"]) + 1)
                ax.plot(it, res["alpha_values_z"], label=f"p={p} Z")
                if fz is not None:
                    ax.plot(it, fz, "--", label=f"p={p} Z fit (R^2={r2z:.3f})")
                if fx is not None:
                    ax.plot(it, res["alpha_values_x"], ":", label=f"p={p} X")
                    ax.plot(it, fx, "-.", label=f"p={p} X fit (R^2={r2x:.3f})")
        if ax is not None:
            ax.set_title(f"n={name}"); ax.set_xlabel("Iteration"); ax.set_ylabel("Alpha"); ax.grid(True, ls="-", alpha=0.4); ax.legend(fontsize=8)
    if plt is not None:
        plt.tight_layout(); plt.savefig(filename, dpi=300); plt.close()
        print(f"\nAlpha linearity plot saved to {filename}")
    return r2_values
