"""Plotting is out of scope for the hot path; this stub keeps
`import visualization` working.  With matplotlib installed the three functions draw basic figures."""


def _plt():
    try:
        import matplotlib.pyplot as plt
        return plt
    except Exception:
        print("matplotlib is not installed; nothing drawn")
        return None


def visualize_grid_and_multiple_paths(grid_data, start_node_pos, target_node_pos, paths_dict, title=""):
    plt = _plt()
    if plt is None:
        return
    plt.figure()
    plt.imshow(grid_data == 1, cmap="gray_r")
    for name, path in paths_dict.items():
        if path:
            plt.plot([c for _, c in path], [r for r, _ in path], label=name)
    plt.legend()
    plt.title(title)


def visualize_pheromone_matrix(grid, pheromone_matrix, title=""):
    plt = _plt()
    if plt is None:
        return
    plt.figure()
    plt.imshow(pheromone_matrix)
    plt.title(title)


def plot_convergence_curve(data, name, color=None):
    plt = _plt()
    if plt is None:
        return
    plt.figure()
    plt.plot([d for d in data if d is not None], color=color)
    plt.title(f"{name} convergence")
