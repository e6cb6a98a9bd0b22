"""Train entry point — same command line as the reference's train.py (flags :45-98):

    torchrun --nnodes=1 --nproc_per_node=4 train.py configs/flux/arcflux_2nfe_k16.py --launcher pytorch --diff_seed
    python train.py configs/flux/arcflux_tiny_smoke.py --gpu-id 0

Config -> `build_model` -> `build_dataset` -> `lakonlab.apis.train_model` (runner DynamicIterBasedRunnerMod). One process
per GPU; `--launcher pytorch` reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the environment (NCCL).
"""
import argparse
import datetime
import logging
import os
import os.path as osp
import time

import torch
import torch.distributed as dist

from lakonlab import __version__
from lakonlab.apis import train_model
from lakonlab.datasets import build_dataset
from lakonlab.models import build_model
from lakonlab.utils import Config, DictAction


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description='Train a model')
    parser.add_argument('config', help='train config file path')
    parser.add_argument('--work-dir', help='the dir to save logs and models')
    parser.add_argument('--resume-from', help='the checkpoint file to resume from')
    parser.add_argument('--no-validate', action='store_true', help='whether not to evaluate the checkpoint during training')
    group_gpus = parser.add_mutually_exclusive_group()
    group_gpus.add_argument('--gpus', type=int, help='(Deprecated, please use --gpu-id)')
    group_gpus.add_argument('--gpu-ids', type=int, nargs='+', help='(Deprecated, please use --gpu-id)')
    group_gpus.add_argument('--gpu-id', type=int, default=0, help='id of gpu to use (non-distributed training)')
    parser.add_argument('--seed', type=int, default=2021, help='random seed')
    parser.add_argument('--diff_seed', action='store_true', help='Whether or not set different seeds for different ranks')
    parser.add_argument('--deterministic', action='store_true', help='accepted for compatibility: the native path is deterministic')
    parser.add_argument('--cfg-options', nargs='+', action=DictAction, help='override config settings, key=value pairs')
    parser.add_argument('--launcher', choices=['none', 'pytorch', 'slurm', 'mpi'], default='none', help='job launcher')
    parser.add_argument('--local-rank', '--local_rank', type=int, default=0)
    args = parser.parse_args(argv)
    if 'LOCAL_RANK' not in os.environ:
        os.environ['LOCAL_RANK'] = str(args.local_rank)
    return args


def main(argv=None):
    args = parse_args(argv)
    cfg = Config.fromfile(args.config)
    if args.cfg_options is not None:
        cfg.merge_from_dict(args.cfg_options)
    if args.work_dir is not None:
        cfg.work_dir = args.work_dir
    elif cfg.get('work_dir', None) is None:
        cfg.work_dir = osp.join('./work_dirs', osp.splitext(osp.basename(args.config))[0])
    if args.resume_from is not None:
        cfg.resume_from = args.resume_from

    if args.launcher == 'none':
        distributed = False
        gpu_id = args.gpu_ids[0] if args.gpu_ids else args.gpu_id
        cfg.gpu_ids = [gpu_id]
        torch.cuda.set_device(gpu_id)
    elif args.launcher == 'pytorch':
        distributed = True
        local = int(os.environ['LOCAL_RANK'])
        torch.cuda.set_device(local)
        dist.init_process_group(backend=(cfg.get('dist_params') or {}).get('backend', 'nccl'),
                                timeout=datetime.timedelta(seconds=3600), device_id=torch.device('cuda', local))
        cfg.gpu_ids = list(range(dist.get_world_size()))
    else:
        raise NotImplementedError(f"launcher '{args.launcher}' is not supported by this build (use 'pytorch')")
    rank = dist.get_rank() if distributed else 0

    os.makedirs(osp.abspath(cfg.work_dir), exist_ok=True)
    timestamp = time.strftime('%Y%m%d_%H%M%S', time.localtime())
    logging.basicConfig(level=getattr(logging, cfg.get('log_level', 'INFO')) if rank == 0 else logging.WARNING,
                        format='%(asctime)s - %(name)s - %(levelname)s - %(message)s',
                        handlers=[logging.StreamHandler()] + ([logging.FileHandler(osp.join(cfg.work_dir, f'{timestamp}.log'))]
                                                              if rank == 0 else []))
    logger = logging.getLogger('lakonlab')
    if rank == 0:
        cfg.dump(osp.join(cfg.work_dir, osp.basename(args.config)))
    logger.info('Distributed training: %s', distributed)
    logger.info('Config:\n%s', cfg.pretty_text)

    # mmgen.apis.set_random_seed(seed, use_rank_shift=diff_seed): rank-shifted seeds give every rank its own noise / prompts
    seed = args.seed + (rank if args.diff_seed else 0)
    logger.info('Set random seed to %d, deterministic: %s, use_rank_shift: %s', args.seed, args.deterministic, args.diff_seed)
    torch.manual_seed(seed)
    cfg.seed = args.seed
    meta = dict(seed=args.seed, exp_name=osp.basename(args.config), lakonlab_version=__version__)

    model = build_model(cfg.model.to_dict() if hasattr(cfg.model, 'to_dict') else dict(cfg.model),
                        train_cfg=dict(cfg.train_cfg), test_cfg=dict(cfg.get('test_cfg') or {}), seed=args.seed)
    model.set_seed(seed)
    datasets = [build_dataset(cfg.data.train)]
    runner = train_model(model, datasets, cfg, distributed=distributed, validate=(not args.no_validate),
                         timestamp=timestamp, meta=meta)
    if distributed:
        dist.destroy_process_group()
    return runner


if __name__ == '__main__':
    main()
