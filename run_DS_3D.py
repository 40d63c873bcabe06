"""Fixed domain decomposition, one model for all subdomains (reference run_DS_3D.py).

    python run_DS_3D.py --mode pred --model neuralop --dataset synthetic \
        --exp_name duct_neuralop --exp_config configs/exp_config/teecnet_duct.yaml
"""
from fesr_b200.cli import main, pred_graph_ALDD, train_graph_ALDD  # noqa: F401

if __name__ == '__main__':
    main(adaptive=False)
