"""Plotting is out of scope for the B200 hot path.
The `plot=` kwargs of manifold2d are accepted; without matplotlib they are
no-ops."""


def _try_pyplot():
    try:
        import matplotlib.pyplot as plt  # noqa: F401
        return plt
    except Exception:
        return None


def plot_img_grid(imgdata, d, **kwargs):
    plt = _try_pyplot()
    if plt is None:
        return
    import torch
    n = imgdata.shape[0]
    h, w = imgdata.shape[-2:]
    canvas = imgdata.reshape(d, n // d, h, w).permute(0, 2, 1, 3).reshape(d * h, (n // d) * w)
    plt.figure(figsize=(8, 8))
    plt.imshow(canvas, cmap=kwargs.get("cmap", "gnuplot"), origin=kwargs.get("origin", "upper"),
               extent=kwargs.get("extent"))
    plt.show()


def plot_spect_grid(spectra, d, **kwargs):
    plt = _try_pyplot()
    if plt is None:
        return
    _, axes = plt.subplots(d, d, figsize=(8, 8))
    for ax, y in zip(axes.flat, spectra):
        ax.plot(y.squeeze())
        ax.set_ylim(*kwargs.get("ylim", (0, 1)))
    plt.show()


def plot_grid_traversal(imgdata, d, **kwargs):
    plot_img_grid(imgdata, d, **kwargs)
