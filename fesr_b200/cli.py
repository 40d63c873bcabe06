"""Shared body of run_DS_3D.py / run_ALDS_3D.py (reference run_DS_3D.py:10-71, run_ALDS_3D.py:10-73).

Same flags (``--mode train|pred``; ``predict`` is accepted as an alias since the reference's README
documents it), same helper names, same "Prediction time" / "Reconstruction time"
prints -- timed with CUDA events instead of unsynchronised wall clock.  Launch with torchrun for
one process per GPU; a single process uses the current device.
"""
from __future__ import annotations

import os

import torch

from .models.scheduler_gnn import GNNPartitionScheduler
from .utils import (init_classifier, init_dataset, init_encoder, init_model, load_yaml, parse_args)


def _init_distributed():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return int(os.environ.get("RANK", "0")), world


def train_graph_ALDD(exp_name, model, dataset, num_partitions, train_config, start_from_pretrained=False, **kwargs):
    scheduler = GNNPartitionScheduler(exp_name, num_partitions, dataset, model, train=True, **kwargs)
    scheduler.train(train_config, start_from_pretrained=start_from_pretrained)
    return scheduler


def pred_graph_ALDD(idxs, exp_name, model, dataset, num_partitions, save_mode, **kwargs):
    kwargs.pop('sub_size', None)
    scheduler = GNNPartitionScheduler(exp_name, num_partitions, dataset, model, train=False, **kwargs)
    rank = int(os.environ.get("RANK", "0"))
    results = []
    for idx in idxs:
        x = dataset.get_one_full_sample(idx)
        t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0.record()
        pred_y_list, ref_y_list, model_idx, weights_list = scheduler.predict(x)
        t1.record()
        pred_y = dataset.reconstruct_from_partition(pred_y_list, ref_y_list, idx, model_idx, weights_list)
        t2.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f'Prediction time: {t0.elapsed_time(t1) / 1e3}')
            print(f'Reconstruction time: {t1.elapsed_time(t2) / 1e3}')
            os.makedirs(f'logs/vtk/{exp_name}', exist_ok=True)
            pred_y.write_vtu(f'logs/vtk/{exp_name}/pred_{idx}.vtu')
            print('Prediction done!')
        results.append(pred_y)
    return results


def main(adaptive: bool):
    args = parse_args()
    rank, world = _init_distributed()
    exp_config = load_yaml(args.exp_config)
    train_config = load_yaml(args.train_config)
    n_clusters = exp_config['n_clusters']
    model = init_model(args.model, **exp_config)
    if args.precision:
        model.precision = args.precision
    dataset = init_dataset(args.dataset, **exp_config)
    extra = {}
    if adaptive:
        extra = {"encoder": init_encoder(args.encoder, **exp_config),
                 "classifier": init_classifier(args.classifier, **exp_config)}
    if rank == 0:
        print('Dataset loaded!')
    if args.mode == 'train':
        train_graph_ALDD(args.exp_name, model, dataset, n_clusters, train_config, **extra)
    elif args.mode in ('pred', 'predict'):
        pred_graph_ALDD(exp_config['idxs'], args.exp_name, model, dataset, n_clusters, 'save_png', **extra)
    else:
        raise SystemExit(f"unknown --mode {args.mode!r} (train | pred)")
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
