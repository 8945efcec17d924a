"""Drop-in for /root/reference/pangenomix/plot.py: mean pan/core curve (+ optional PNG).

Only the column means (plot.py:8-11) are on the hot path -- they feed
``fit_heaps_by_iteration``.  The figure (plot.py:21-41) is drawn only when matplotlib is
importable and a file stem is given; the returned DataFrame is the same either way.
"""
import pandas as pd


def calculate_mean(df_pan_core, jpgName=None):
    mean_df = pd.DataFrame([df_pan_core.mean()], columns=df_pan_core.columns)
    if jpgName is not None:
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except ImportError:
            print("matplotlib is not installed: skipping %s_plot.png" % jpgName)
            return mean_df
        half = mean_df.shape[1] // 2
        x = range(1, half + 1)
        plt.figure()
        plt.plot(x, mean_df.iloc[0, :half].values, label='Pangenome size')
        plt.plot(x, mean_df.iloc[0, half:].values, label='Core gene size')
        plt.xlabel('number of genomes')
        plt.ylabel('number of genes')
        plt.legend()
        plt.savefig('%s_plot.png' % jpgName)
        plt.close()
    return mean_df
