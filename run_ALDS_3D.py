"""Adaptive decomposition: every subdomain is routed by PCA + k-means to one of n_clusters
per-cluster models (reference run_ALDS_3D.py).

    torchrun --nproc-per-node 8 run_ALDS_3D.py --mode pred --model neuralop --dataset synthetic ...
"""
from fesr_b200.cli import main, pred_graph_ALDD, train_graph_ALDD  # noqa: F401

if __name__ == '__main__':
    main(adaptive=True)
