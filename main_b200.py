#!/usr/bin/env python
"""main_b200.py -- the reference's training driver (main.py:79-233) on this implementation.

Same command line as the reference's main.py (README.md:5-32), same one_iter flow
(zero_grad -> forward -> loss -> backward -> quantizer.record(user) for every user, then
quantizer.apply() and optimizer.step(), main.py:216-233), same learning-rate schedule and SignSGD
overrides (main.py:136-157) -- but the compressors / quantizers are the B200 ones, and the
simulated users can be REAL ranks:

    python main_b200.py --network fcn --dataset synthetic --quantizer hsq --c-dim 16 --k-bit 8 --n-bit 6 \
        --num-users 8 --logdir runs/hsq                          # 8 simulated users in one process
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 main_b200.py \
        --num-users 8 ...                                        # one user per GPU, packed codes over NVLink

Differences from the reference, all forced by the environment: the TensorFlow-1 `Logger`
(logger.py:15-20) is replaced by a CSV writer with the same scalar_summary(tag, value, step) call;
`--dataset synthetic` (default; there is no network to download MNIST / CIFAR) draws a fixed,
learnable classification set of the chosen dataset's shape; `--network` knows the reference's FCN
(models/fcn.py) natively and imports any other reference model class from `--models-path`
(e.g. /root/reference) instead of copying it; `--max-iters` / `--max-epochs` bound a run.
"""
import argparse
import csv
import importlib
import os
import sys
import time
from datetime import datetime

import torch
import torch.nn as nn
import torch.optim as optim

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOSS_FUNC = nn.CrossEntropyLoss()
DATASET_SHAPES = {'mnist': ((1, 28, 28), 10, 60000, 10000), 'cifar10': ((3, 32, 32), 10, 50000, 10000),
                  'cifar100': ((3, 32, 32), 100, 50000, 10000), 'stl10': ((3, 96, 96), 10, 5000, 8000),
                  'svhn': ((3, 32, 32), 10, 73257, 26032), 'tinyimg': ((3, 64, 64), 200, 100000, 10000)}


class Logger(object):
    """TF-free stand-in for the reference's Logger (logger.py:15-20): same scalar_summary call,
    rows `tag,step,value,wall_time` appended to <logdir>/scalars.csv (what converter.py extracts)."""

    def __init__(self, log_dir):
        os.makedirs(log_dir, exist_ok=True)
        self.path = os.path.join(log_dir, "scalars.csv")
        self.fh = open(self.path, "a", newline="")
        self.writer = csv.writer(self.fh)

    def scalar_summary(self, tag, value, step):
        self.writer.writerow([tag, int(step), float(value), time.time()])
        self.fh.flush()


class FCN(nn.Module):
    """The reference's two-layer perceptron (models/fcn.py:5-27): 784 -> 256 -> classes."""

    def __init__(self, D_in=784, H=256, num_classes=10):
        super(FCN, self).__init__()
        self.linear1 = nn.Linear(D_in, H)
        self.linear2 = nn.Linear(H, num_classes)

    def forward(self, x):
        x = x.view(x.shape[0], -1)
        return self.linear2(self.linear1(x).clamp(min=0))


class SyntheticSet(torch.utils.data.Dataset):
    """A fixed, learnable classification set of a dataset's shape: class prototype + noise."""

    def __init__(self, shape, classes, n, seed):
        g = torch.Generator().manual_seed(seed)
        protos = torch.randn(classes, *shape, generator=torch.Generator().manual_seed(4242))
        self.targets = torch.randint(0, classes, (n,), generator=g)
        self.data = protos[self.targets] * 0.5 + torch.randn(n, *shape, generator=g)

    def __len__(self):
        return self.data.shape[0]

    def __getitem__(self, i):
        return self.data[i], self.targets[i]


def make_loaders(args):
    """global batch = batch_size * num_users, like dataloaders.py:12,40."""
    name = args.dataset
    shape_of = 'mnist' if name == 'synthetic' else name
    shape, classes, n_train, n_test = DATASET_SHAPES[shape_of]
    if args.network == 'fcn' and name == 'synthetic':
        shape, classes = (1, 28, 28), 10
    args.num_classes = classes
    if name != 'synthetic':
        raise SystemExit("--dataset %s needs the torchvision download of the reference's dataloaders.py; this "
                         "environment has no network: use --dataset synthetic" % name)
    n_train, n_test = min(n_train, args.train_size), min(n_test, args.test_size)
    train = SyntheticSet(shape, classes, n_train, seed=args.seed)
    test = SyntheticSet(shape, classes, n_test, seed=args.seed + 1)
    g = torch.Generator().manual_seed(args.seed)
    train_loader = torch.utils.data.DataLoader(train, batch_size=args.batch_size * args.num_users, shuffle=True, generator=g)
    test_loader = torch.utils.data.DataLoader(test, batch_size=args.test_batch_size, shuffle=False)
    return train_loader, test_loader


def make_model(args):
    if args.network == 'fcn':
        return FCN(num_classes=args.num_classes)
    if not args.models_path:
        raise SystemExit("--network %s: give --models-path <reference checkout> to import the reference's model "
                         "classes (they are not part of the compression path and are not rebuilt here)" % args.network)
    sys.path.insert(0, args.models_path)
    models = importlib.import_module("models")
    table = {'resnet18': 'ResNet18', 'resnet34': 'ResNet34', 'resnet50': 'ResNet50', 'resnet101': 'ResNet101',
             'resnet152': 'ResNet152', 'vgg11': 'vgg11', 'vgg13': 'vgg13', 'vgg16': 'vgg16', 'vgg19': 'vgg19',
             'dense': 'densenet_cifar'}
    return getattr(models, table[args.network])(num_classes=args.num_classes)


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description='Gradient Quantization Samples (B200)')
    parser.add_argument('--network', type=str, default='fcn',
                        choices=['resnet18', 'resnet34', 'resnet50', 'resnet101', 'resnet152', 'vgg11', 'vgg13', 'vgg16',
                                 'vgg19', 'dense', 'fcn'])
    parser.add_argument('--dataset', type=str, default='synthetic', choices=['synthetic'] + sorted(DATASET_SHAPES))
    parser.add_argument('--num-classes', type=int, default=10)
    parser.add_argument('--quantizer', type=str, default='hsq', choices=['sgd', 'qsgd', 'hsq', 'sign', 'topk'])
    parser.add_argument('--mode', type=str, default='ps', choices=['ps', 'ring'])
    parser.add_argument('--scale', type=str, default="exp")
    parser.add_argument('--c-dim', type=int, default=32)
    parser.add_argument('--k-bit', type=int, default=8)
    parser.add_argument('--n-bit', type=int, default=8)
    parser.add_argument('--cr', type=int, default=256)
    parser.add_argument('--random', type=int, default=True)
    parser.add_argument('--num-users', type=int, default=8, metavar='N')
    parser.add_argument('--logdir', type=str, default=None)
    parser.add_argument('--batch-size', type=int, default=32, metavar='N')
    parser.add_argument('--test-batch-size', type=int, default=1000, metavar='N')
    parser.add_argument('--epochs', type=int, default=350, metavar='N')
    parser.add_argument('--momentum', type=float, default=0.9, metavar='M')
    parser.add_argument('--weight-decay', type=float, default=5e-4, metavar='M')
    parser.add_argument('--no-cuda', action='store_true', default=False)
    parser.add_argument('--ef', action='store_true', default=False)
    parser.add_argument('--seed', type=int, default=1, metavar='S')
    parser.add_argument('--log-epoch', type=int, default=1, metavar='N')
    parser.add_argument('--save-model', action='store_true', default=False)
    parser.add_argument('--two-phase', action='store_true', default=False)
    # additions (see the module docstring)
    parser.add_argument('--models-path', type=str, default=None)
    parser.add_argument('--max-iters', type=int, default=0, help='stop after this many iterations (0: no limit)')
    parser.add_argument('--max-epochs', type=int, default=0)
    parser.add_argument('--train-size', type=int, default=1 << 30)
    parser.add_argument('--test-size', type=int, default=1 << 30)
    parser.add_argument('--rng', type=str, default='philox', choices=['philox', 'torch'],
                        help="'torch': the reference's CPU uniform stream (per-parameter path), for lock-step comparisons")
    return parser.parse_args(argv)


def one_iter(model, device, loss_func, optimizer, quantizer, train_data, users, epoch):
    """main.py:216-233; `users` = the simulated users this process plays (all of them, or its rank)."""
    model.train()
    all_losses = []
    for user_id in users:
        optimizer.zero_grad()
        data, target = train_data[user_id]
        data, target = data.to(device), target.to(device)
        loss = loss_func(model(data), target)
        all_losses.append(loss)
        loss.backward()
        quantizer.record(user_id, epoch=epoch)
    quantizer.apply()
    optimizer.step()
    return torch.stack(all_losses).mean()


def test(args, model, device, test_loader, quiet=False):
    model.eval()
    test_loss, correct = 0.0, 0
    with torch.no_grad():
        for data, target in test_loader:
            data, target = data.to(device), target.to(device)
            output = model(data)
            test_loss += LOSS_FUNC(output, target).sum().item()
            correct += output.argmax(dim=1, keepdim=True).eq(target.view_as(output.argmax(dim=1, keepdim=True))).sum().item()
    n = len(test_loader.dataset)
    if not quiet:
        print('\nTest set: Average loss: {:.4f}, Accuracy: {}/{} ({:.2f}%)\n'.format(test_loss / n, correct, n, 100. * correct / n))
    return correct / n


def train(args, model, device, train_loader, test_loader, optimizer, quantizer, epoch, logger, state):
    batch_size, num_users = args.batch_size, args.num_users
    n = len(train_loader.dataset)
    iteration = n // (num_users * batch_size) + int(n % (num_users * batch_size) != 0)
    log_interval = [iteration // args.log_epoch * (i + 1) for i in range(args.log_epoch)]
    users = [state["rank"]] if state["world"] > 1 else list(range(num_users))
    loss = torch.zeros(())
    for batch_idx, (data, target) in enumerate(train_loader):
        ub = len(data) // num_users
        train_data = [(data[u * ub:(u + 1) * ub], target[u * ub:(u + 1) * ub]) for u in range(num_users - 1)]
        train_data.append((data[(num_users - 1) * ub:], target[(num_users - 1) * ub:]))
        loss = one_iter(model, device, LOSS_FUNC, optimizer, quantizer, train_data, users, epoch=epoch)
        state["iters"] += 1
        if state["world"] > 1:   # the reference reports the mean loss over users
            t = loss.detach().clone()
            torch.distributed.all_reduce(t)
            loss = t / state["world"]
        state["losses"].append(loss.item())
        if (batch_idx + 1) in log_interval:
            acc = test(args, model, device, test_loader, quiet=state["rank"] != 0)
            if state["rank"] == 0:
                print('Train Epoch: {} [{}/{} ({:.0f}%)]\tLoss: {:.6f}\t Test Accuracy: {:.2f}%'.format(
                    epoch, batch_idx * num_users * batch_size + len(data), n, 100. * batch_idx / len(train_loader),
                    loss.item(), acc * 100))
                for tag, value in {'loss': loss.item(), 'accuracy(%)': acc * 100}.items():
                    logger.scalar_summary(tag, value, iteration * (epoch - 1) + batch_idx)
        if args.max_iters and state["iters"] >= args.max_iters:
            return True
    if state["rank"] == 0:
        print('Train Epoch: {} Done.\tLoss: {:.6f}'.format(epoch, loss.item()))
    return False


def main(argv=None):
    args = parse_args(argv)
    import gq_b200
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.no_cuda or not torch.cuda.is_available():
        raise SystemExit("main_b200.py is CUDA-only (the compressors have no CPU path); run the reference for --no-cuda")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        if world != args.num_users:
            raise SystemExit("one user per rank: --num-users %d != WORLD_SIZE %d" % (args.num_users, world))
        torch.distributed.init_process_group("nccl", device_id=device)
    if args.logdir is None:
        assert False, "The logdir is not defined"
    logger = Logger(args.logdir) if rank == 0 else None
    compressor = {'sgd': gq_b200.IdenticalCompressor, 'qsgd': gq_b200.QSGDCompressor,
                  'hsq': gq_b200.NearestNeighborCompressor, 'sign': gq_b200.SignSGDCompressor,
                  'topk': gq_b200.TopKSparsificationCompressor}[args.quantizer]
    torch.manual_seed(args.seed)            # identical model replicas and data order on every rank (main.py:127)
    train_loader, test_loader = make_loaders(args)
    model = make_model(args).to(device)
    if args.rng == 'torch':
        args.fused = False                  # the reference's CPU draws exist only on the per-parameter path
    quantizer = gq_b200.Quantizer(compressor, model.parameters(), args)
    optimizer = optim.SGD(model.parameters(), lr=0.1, momentum=args.momentum, weight_decay=args.weight_decay)
    if args.dataset in ('mnist', 'synthetic'):
        epochs, lrs, args.epochs = [], [], 20
    elif args.dataset == 'tinyimg':
        epochs, lrs, args.epochs = [51], [0.01], 1000
    else:
        epochs, lrs, args.epochs = [51, 71], [0.01, 0.005], 150
    if args.quantizer == 'sign':            # main.py:149-157
        epochs, lrs, args.epochs = [51, 71], [0.0005, 0.0001], 150
        args.momentum, args.weight_decay = 0.0, 0.1
        optimizer = optim.SGD(model.parameters(), lr=1e-3, momentum=args.momentum, weight_decay=args.weight_decay)
    if args.max_epochs:
        args.epochs = min(args.epochs, args.max_epochs - 1)
    state = {"rank": rank, "world": world, "iters": 0, "losses": []}
    t0 = time.time()
    for epoch in range(1, args.epochs + 2):
        for i_epoch, i_lr in zip(epochs, lrs):
            if epoch == i_epoch:
                optimizer = optim.SGD(model.parameters(), lr=i_lr, momentum=args.momentum, weight_decay=5e-4)
        if train(args, model, device, train_loader, test_loader, optimizer, quantizer, epoch, logger, state):
            break
    torch.cuda.synchronize()
    if rank == 0:
        acc = test(args, model, device, test_loader, quiet=True)
        print("done: %d iterations in %.1f s, final loss %.6f, test accuracy %.2f%%"
              % (state["iters"], time.time() - t0, state["losses"][-1], acc * 100))
        if args.save_model:
            torch.save(model.state_dict(), "saved_{}_{}.pt".format(args.network, datetime.now()))
    if world > 1:
        torch.distributed.destroy_process_group()
    return state


if __name__ == "__main__":
    main()
