"""Mirror of the reference's trainers/base_trainer.py:14-157 for the B200 path.

Differences that are deliberate and documented:
  * the device must be CUDA — `cfg.device` values 'cpu'/'cuda' are accepted like the reference does
    (base_trainer.py:20-25) but 'cpu' raises: this framework has no CPU path;
  * `_optimizer` returns a FusedOptimizer descriptor (kind / lr / weight_decay / step / moments) consumed by the
    fused kernels instead of a torch.optim object; names and error behaviour follow base_trainer.py:34-43.
"""
from __future__ import annotations

import logging
import os
from abc import ABC, abstractmethod

import torch

from .. import _cabi

try:  # the reference logs with loguru; stay importable without it
    from loguru import logger
except Exception:  # pragma: no cover
    logger = logging.getLogger("yelprecommendation_b200")


class FusedOptimizer:
    """Hyper-parameters + state of torch.optim.{SGD,Adam,AdamW} with default betas/eps, applied by the kernels."""

    def __init__(self, name: str, lr: float, weight_decay: float = 0.0):
        self.name = str(name).lower()
        if self.name not in _cabi.OPT_KINDS:
            logger.error(f"Optimizer Not Exists: {name}")
            raise NotImplementedError(f"Optimizer Not Exists: {name}")
        self.lr, self.weight_decay = float(lr), float(weight_decay)
        self.step_count = 0          # optimizer steps taken so far
        self.state = {}              # name -> (exp_avg, exp_avg_sq)

    @property
    def needs_moments(self) -> bool:
        return self.name != "sgd"

    def opt_struct(self, next_step: int) -> _cabi.YrOpt:
        return _cabi.make_opt(self.name, self.lr, self.weight_decay, next_step)

    def zero_grad(self):  # API parity with torch.optim; gradients never materialise here
        return None


class BaseTrainer(ABC):
    def __init__(self, cfg) -> None:
        self.cfg = cfg
        self.device: torch.device = self._device(self.cfg.device)
        os.makedirs(self.cfg.model_dir, exist_ok=True)

    def _device(self, device_name: str) -> torch.device:
        name = str(device_name).lower()
        if name.startswith("cuda"):
            if not torch.cuda.is_available():
                raise _cabi.YelprecError("cfg.device='cuda' but no CUDA device is visible (no CPU fallback)")
            return torch.device("cuda", torch.cuda.current_device())
        if name == "cpu":
            raise _cabi.YelprecError("yelprecommendation_b200 is the B200 path: set cfg.device='cuda' "
                                     "(the CPU path is the reference itself)")
        logger.error(f"Not supported device: {device_name}")
        raise _cabi.YelprecError(f"Not supported device: {device_name}")

    def _optimizer(self, optimizer_name: str, model, learning_rate: float, weight_decay: float = 0) -> FusedOptimizer:
        return FusedOptimizer(optimizer_name, learning_rate, weight_decay)

    def _is_surpass_best_metric(self, **metric) -> bool:
        (valid_loss, valid_precision, valid_recall, valid_map, valid_ndcg) = metric["current"]
        (best_loss, best_precision, best_recall, best_map, best_ndcg) = metric["best"]
        key = self.cfg.best_metric
        if key == "loss":
            return valid_loss < best_loss
        if key == "precision":
            return valid_precision > best_precision
        if key == "recall":
            return valid_recall > best_recall
        if key == "map":
            return valid_map > best_map
        if key == "ndcg":
            return valid_ndcg > best_ndcg
        return False

    def run(self, train_dataloader, valid_dataloader, valid_eval_data):
        """Epoch loop with early stopping and best-model checkpointing (mf_trainer.py:34-97)."""
        logger.info("[Trainer] run...")
        best = (1e+6, .0, .0, .0, .0)
        endurance = 0
        for epoch in range(self.cfg.epochs):
            train_loss = self.train(train_dataloader)
            valid_loss = self.validate(valid_dataloader)
            p, r, m, n = self.evaluate(valid_eval_data, "valid")
            logger.info(f"[Trainer] epoch: {epoch} > train loss: {train_loss:.4f} / valid loss: {valid_loss:.4f} / "
                        f"precision@K : {p:.4f} / Recall@K: {r:.4f} / MAP@K: {m:.4f} / NDCG@K: {n:.4f}")
            if getattr(self.cfg, "wandb", False):
                import wandb
                wandb.log({"train_loss": train_loss, "valid_loss": valid_loss, "valid_Precision@K": p,
                           "valid_Recall@K": r, "valid_MAP@K": m, "valid_NDCG@K": n})
            current = (valid_loss, p, r, m, n)
            if self._is_surpass_best_metric(current=current, best=best):
                logger.info("[Trainer] update best model...")
                best = current
                endurance = 0
                torch.save(self.model.state_dict(), f"{self.cfg.model_dir}/best_model.pt")
            else:
                endurance += 1
                if endurance > self.cfg.patience:
                    logger.info("[Trainer] ealry stopping...")
                    break

    @abstractmethod
    def train(self, train_dataloader) -> float:
        ...

    @abstractmethod
    def validate(self, valid_dataloader) -> float:
        ...

    @abstractmethod
    def evaluate(self, eval_data, mode="valid") -> tuple:
        ...

    def load_best_model(self):
        logger.info("[Trainer] Load best model...")
        self.model.load_state_dict(torch.load(f"{self.cfg.model_dir}/best_model.pt"))


class BatchStager:
    """Collates DataLoader batches ({'user_id','pos_item','neg_item'} int64 CPU tensors) into pinned staging
    buffers and ships `chunk` batches per host->device copy, so the persistent kernels see many steps per launch."""

    def __init__(self, device, batch_cap: int, chunk: int):
        self.device, self.cap, self.chunk = device, int(batch_cap), int(chunk)
        n = self.cap * self.chunk
        self.host = [torch.empty(3, n, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.dev = [torch.empty(3, n, dtype=torch.int64, device=device) for _ in range(2)]
        self.events = [None, None]
        self.which = 0

    def chunks(self, dataloader):
        """Yields (uid, pos, neg, n_triples, B) device views; every batch in a chunk has size B except the last.
        Batches are collected by reference and collated with three torch.cat(out=pinned) calls per chunk, so the
        per-batch host cost is a few attribute lookups, not three small copies."""
        us, ps, ns = [], [], []
        fill, B = 0, None
        cap_n = self.cap * self.chunk

        def flush():
            nonlocal us, ps, ns, fill, B
            host = self._acquire()
            torch.cat(us, out=host[0, :fill])
            torch.cat(ps, out=host[1, :fill])
            torch.cat(ns, out=host[2, :fill])
            out = self._ship(host, fill, B)
            us, ps, ns, fill, B = [], [], [], 0, None
            return out

        for data in dataloader:
            u, p, n = data["user_id"], data["pos_item"], data["neg_item"]
            nb = int(u.numel())
            if nb == 0:
                continue
            short_pending = B is not None and fill % B != 0      # a short batch must end its chunk
            if B is not None and (nb > B or short_pending or fill + nb > cap_n or nb > self.cap):
                yield flush()
            if nb > cap_n:
                raise _cabi.YelprecError(f"batch of {nb} triples exceeds the staging capacity {cap_n}")
            if B is None:
                B = nb
            if u.device.type != "cpu" or u.dtype != torch.int64 or p.dtype != torch.int64 or n.dtype != torch.int64:
                u, p, n = (t.to("cpu", torch.int64) for t in (u, p, n))
            us.append(u.reshape(-1)); ps.append(p.reshape(-1)); ns.append(n.reshape(-1))
            fill += nb
            if fill + B > cap_n or nb < B:
                yield flush()
        if fill:
            yield flush()

    def _acquire(self):
        self.which ^= 1
        ev = self.events[self.which]
        if ev is not None:
            ev.synchronize()          # the copy out of this staging buffer has finished
        return self.host[self.which]

    def _ship(self, host, fill, B):
        dev = self.dev[self.which]
        dev[:, :fill].copy_(host[:, :fill], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.events[self.which] = ev
        self.h2d_bytes = getattr(self, "h2d_bytes", 0) + 3 * 8 * fill
        return dev[0, :fill], dev[1, :fill], dev[2, :fill], fill, B
