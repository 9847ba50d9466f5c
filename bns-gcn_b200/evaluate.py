"""Evaluation + checkpoint branch of the reference's ``run`` (train.py:14-61, 308-321, 354-356, 427-456).

Off the throughput path (the reference's own benchmark runs use ``--no-eval``, README.md:110), kept so that a user of
the reference finds the same behaviour: every ``log_every`` epochs rank 0 saves ``model.state_dict()`` to
``checkpoint/<graph_name>_p<rate>_<epoch>.pth.tar``, evaluates on the full (validation) graph, appends the line to
``results/<dataset>_n<parts>_p<rate>.txt``, keeps the best model, and at the end writes ``<graph_name>_final.pth.tar``
and prints the test accuracy.  State-dict keys are the reference's parameter names (the module mirrors keep them), so
checkpoints are interchangeable.

Differences, deliberate: the reference copies the model to the CPU and evaluates in a thread pool with DGL on the host;
here the copy stays on the GPU and the full-graph forward uses the same SpMM / dense kernels as training
(``FullGraphHandle``: module/layer.py:39-45, 93-102 eval branches), synchronously.
"""
from __future__ import annotations

import copy
import dataclasses
import os
from typing import Dict, Optional

import torch

from .data.partition import induced_subgraph
from .data.synthetic import FullGraph


def calc_acc(logits: torch.Tensor, labels: torch.Tensor) -> float:
    """train.py:14-20: accuracy for single-label tasks, micro-F1 of ``logits > 0`` for multi-label ones
    (``sklearn.metrics.f1_score(labels, logits > 0, average='micro')`` = 2 TP / (2 TP + FP + FN))."""
    if labels.dim() == 1:
        if labels.shape[0] == 0:
            return 0.0
        return (logits.argmax(dim=1) == labels).sum().item() / labels.shape[0]
    pred, lab = logits > 0, labels > 0.5
    tp = (pred & lab).sum().item()
    fp = (pred & ~lab).sum().item()
    fn = (~pred & lab).sum().item()
    den = 2 * tp + fp + fn
    return 2.0 * tp / den if den else 0.0


@dataclasses.dataclass
class EvalGraph:
    """The ``val_g`` / ``test_g`` of the reference: a full homogeneous graph with its node data on the device."""
    handle: object                      # FullGraphHandle
    ndata: Dict[str, torch.Tensor]      # feat, label, train_mask, val_mask, test_mask


def build_eval_graph(fg: FullGraph, device) -> EvalGraph:
    from . import ops
    from .graph import FullGraphHandle
    dev = torch.device(device)
    a = ops.DeviceGraph.from_csr(fg.indptr.to(dev), fg.src.to(torch.int32).to(dev), fg.n_nodes)
    handle = FullGraphHandle(a, fg.in_degrees().to(dev), fg.out_degrees().to(dev))
    nd = {"feat": fg.feat.to(dev), "label": fg.label.to(dev), "train_mask": fg.train_mask.to(dev),
          "val_mask": fg.val_mask.to(dev), "test_mask": fg.test_mask.to(dev)}
    return EvalGraph(handle, nd)


def eval_graphs(fg: FullGraph, inductive: bool, device):
    """train.py:313-321: transductive -> the full graph for both; inductive -> (train | val) subgraph and the full graph
    (helper/utils.py:226-230)."""
    if not inductive:
        g = build_eval_graph(fg, device)
        return g, g
    return build_eval_graph(induced_subgraph(fg, fg.train_mask | fg.val_mask), device), build_eval_graph(fg, device)


def _emit(buf: str, result_file_name: Optional[str]) -> None:
    if result_file_name is not None:
        with open(result_file_name, 'a+') as f:
            f.write(buf + '\n')
    print(buf)


@torch.no_grad()
def evaluate_induc(name, model, g: EvalGraph, mode, result_file_name=None):
    """train.py:22-41.  ``mode``: 'val' or 'test'."""
    model.eval()
    feat, labels = g.ndata['feat'], g.ndata['label']
    mask = g.ndata[mode + '_mask']
    logits = model(g.handle, feat)
    acc = calc_acc(logits[mask], labels[mask])
    _emit("{:s} | Accuracy {:.2%}".format(name, acc), result_file_name)
    return model, acc


@torch.no_grad()
def evaluate_trans(name, model, g: EvalGraph, result_file_name=None):
    """train.py:44-61."""
    model.eval()
    feat, labels = g.ndata['feat'], g.ndata['label']
    val_mask, test_mask = g.ndata['val_mask'], g.ndata['test_mask']
    logits = model(g.handle, feat)
    val_acc = calc_acc(logits[val_mask], labels[val_mask])
    test_acc = calc_acc(logits[test_mask], labels[test_mask])
    _emit("{:s} | Validation Accuracy {:.2%} | Test Accuracy {:.2%}".format(name, val_acc, test_acc), result_file_name)
    return model, val_acc


def result_file_name(args) -> str:
    """train.py:356."""
    return 'results/%s_n%d_p%.2f.txt' % (args.dataset, args.n_partitions, args.sampling_rate)


def checkpoint_path(args, epoch: Optional[int] = None) -> str:
    """train.py:428 (periodic) and :452 (final)."""
    if epoch is None:
        return 'checkpoint/' + args.graph_name + '_final.pth.tar'
    return 'checkpoint/%s_p%.2f_%d.pth.tar' % (args.graph_name, args.sampling_rate, epoch)


def save_checkpoint(model: torch.nn.Module, path: str) -> None:
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, path)


def load_checkpoint(model: torch.nn.Module, path: str, strict: bool = True):
    """Load a checkpoint written by this build or by the reference (same keys, same shapes)."""
    return model.load_state_dict(torch.load(path, map_location='cpu'), strict=strict)


class Evaluator:
    """Rank 0's bookkeeping of train.py:354-356, 427-456."""

    def __init__(self, args, fg: FullGraph, device):
        self.args = args
        os.makedirs('checkpoint/', exist_ok=True)            # train.py:310-311
        os.makedirs('results/', exist_ok=True)
        self.val_g, self.test_g = eval_graphs(fg, args.inductive, device)
        self.best_model, self.best_acc = None, 0.0
        self.result_file_name = result_file_name(args)

    def after_epoch(self, model: torch.nn.Module, epoch: int) -> float:
        """train.py:427-442 at an epoch with ``(epoch + 1) % log_every == 0``."""
        save_checkpoint(model, checkpoint_path(self.args, epoch))
        snap = copy.deepcopy(model)
        was_training = model.training
        if not self.args.inductive:
            _, val_acc = evaluate_trans('Epoch %05d' % epoch, snap, self.val_g, self.result_file_name)
        else:
            _, val_acc = evaluate_induc('Epoch %05d' % epoch, snap, self.val_g, 'val', self.result_file_name)
        if val_acc > self.best_acc or self.best_model is None:
            self.best_acc, self.best_model = val_acc, snap
        model.train(was_training)
        return val_acc

    def finish(self, model: torch.nn.Module) -> float:
        """train.py:446-456."""
        if self.best_model is None:
            self.best_model = copy.deepcopy(model)
        save_checkpoint(self.best_model, checkpoint_path(self.args))
        print('model saved')
        print("Max Validation Accuracy {:.2%}".format(self.best_acc))
        _, acc = evaluate_induc('Test Result', self.best_model, self.test_g, 'test')
        return acc
