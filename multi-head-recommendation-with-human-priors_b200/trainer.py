"""Train / eval loop harness around the B200 hot path: the subset of the reference
`Trainer` (code/REC/trainer/trainer.py) that drives the HSTU model — `fit` (:455-696), `evaluate`
(:826-1153), the LR schedules (utils/lr_scheduler.py:44-116), validation-based early stopping
(utils/utils.py:60-101), checkpoints carrying the reference's state-dict keys (:319-369).

Not a port of the reference trainer: no Lightning / DeepSpeed / wandb / tensorboard.  One process per GPU;
on one GPU the step is a CUDA-graph replay (`GraphedTrainStep`), with a row-sharded table it is the
pre / graph / post step of `GraphedShardedStep`.  Loaders are plain iterables of reference-shaped batches:
  train: (items i64[B, L+P], neg_items i64[B, sets, n], attn_mask i64[B, L+P], tags i64[B, L+P, C] or empty)
  eval : dict(item_seq, item_target, history_index=(u, i), positive_u, target_tags)   (evalset.py:81-155)
"""
import math
import os
import time

import torch
import torch.distributed as dist

from .evaluator import Collector, Evaluator
from .optim import FusedAdamW


# ---- LR schedules (multipliers of the base lr; utils/lr_scheduler.py) -------------------------------------
def constant_schedule(step, warmup_steps=0, total_steps=0):
    return 1.0


def linear_schedule_with_warmup(step, warmup_steps, total_steps, lr_end=1e-7):
    """lr_scheduler.py:44-76 (the multiplier never drops below lr_end = 1e-7)."""
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    return max(lr_end, float(total_steps - step) / float(max(1, total_steps - warmup_steps)))


def cosine_schedule_with_warmup(step, warmup_steps, total_steps, num_cycles=0.5):
    """lr_scheduler.py:79-116."""
    if step < warmup_steps:
        return float(step) / float(max(1, warmup_steps))
    progress = float(step - warmup_steps) / float(max(1, total_steps - warmup_steps))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * progress)))


SCHEDULES = {"cosine": cosine_schedule_with_warmup, "linear": linear_schedule_with_warmup}


def get_model(model_name):
    """REC/utils/utils.py:38-57: module `<model_name.lower()>` of the package, attribute `<model_name>` (HSTU, ComiRec,
    REMI); the same ValueError for a name without a module."""
    import importlib
    import importlib.util
    module_path = f"{__package__}.{model_name.lower()}"
    if importlib.util.find_spec(module_path) is None:
        raise ValueError("`model_name` [{}] is not the name of an existing model.".format(model_name))
    return getattr(importlib.import_module(module_path), model_name)


def early_stopping(value, best, cur_step, max_step, bigger=True):
    """utils/utils.py:60-101 -> (best, cur_step, stop_flag, update_flag)."""
    stop_flag = update_flag = False
    better = value >= best if bigger else value <= best
    if better:
        cur_step, best, update_flag = 0, value, True
    else:
        cur_step += 1
        if cur_step > max_step:
            stop_flag = True
    return best, cur_step, stop_flag, update_flag


def calculate_valid_score(valid_result, eval_pred_len, valid_metric=None):
    """utils/utils.py:104-122: the chosen metric of the last prediction offset."""
    key = f"pred_{eval_pred_len - 1}"
    res = valid_result[key] if key in valid_result else valid_result[sorted(valid_result)[-1]]
    if valid_metric:
        for k, v in res.items():
            if k.lower() == valid_metric.lower():
                return v
        raise KeyError(f"valid_metric {valid_metric} not in {list(res)}")
    return res["recall@10"]


class Trainer(object):
    def __init__(self, config, model, optimizer=None, use_graph=True, log=None):
        self.config, self.model = config, model
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        oa = config.get("optim_args", None) or {}
        self.base_lr = float(oa.get("learning_rate", config.get("learning_rate", 1e-3) or 1e-3))
        wd = float(oa.get("weight_decay", config.get("weight_decay", 0.0) or 0.0))
        self.sharded = getattr(model, "sharded_table", None) is not None
        self.use_graph = bool(use_graph) and next(model.parameters()).is_cuda
        if self.world > 1 and not self.sharded:
            # replicated table on several ranks: the captured single-GPU step contains no gradient exchange (and would
            # capture the negative-id all-gather); only the eager step synchronises gradients (DataParallel)
            self.use_graph = False
        if optimizer is None:
            model.sparse_embedding_grad = True
            optimizer = FusedAdamW(model, lr=self.base_lr, weight_decay=wd,
                                   device_step=self.use_graph and not self.sharded, lazy_table=True)
        self.optimizer = optimizer
        self.scheduler_config = config.get("scheduler_args", None)
        self.total_iters = int(config.get("total_iters", 0) or 0)
        self.eval_interval = int(config.get("eval_freq", 0) or 0)
        self.stopping_step = int(config.get("stopping_step", 10) or 10)
        self.valid_metric = (config.get("valid_metric", "recall@10") or "recall@10").lower()
        self.valid_metric_bigger = bool(config.get("valid_metric_bigger", True))
        self.metric_decimal_place = int(config.get("metric_decimal_place", 7) or 7)
        self.metrics_pred_len_list = list(config["metrics_pred_len_list"])
        self.eval_pred_len = config["eval_pred_len"]
        self.checkpoint_dir = config.get("checkpoint_dir", None)
        self.train_step = 0
        self.best_valid_score = -float("inf") if self.valid_metric_bigger else float("inf")
        self.best_valid_result = None
        self.no_improve_times = 0
        self.train_loss_dict = {}
        self.results = []                      # rows of the results table (trainer.py:645-648)
        self._stepper = None
        self._log = log or (lambda msg: None)
        self.item_feature = None

    # ------------------------------------------------------------------ schedule
    def lr_at(self, step):
        if not self.scheduler_config:
            return self.base_lr
        fn = SCHEDULES.get(self.scheduler_config.get("type"), constant_schedule)
        warm = self.total_iters * self.scheduler_config.get("warmup", 0.001)       # trainer.py:457-458
        return self.base_lr * fn(step, warm, self.total_iters)

    # ------------------------------------------------------------------ one optimisation step
    def _step(self, batch):
        model, opt = self.model, self.optimizer
        dev = next(model.parameters()).device
        Lc = model.max_seq_length
        if self.use_graph:
            n_tok = int(batch[2][:, :Lc].sum())             # host metadata of the collate fn when the batch is on the host
            batch = tuple(t.to(dev, non_blocking=True) for t in batch)
            if self._stepper is None:
                from .graphed import GraphedTrainStep, GraphedShardedStep
                self._stepper = (GraphedShardedStep if self.sharded else GraphedTrainStep)(model, opt, batch)
            return self._stepper(batch, n_tok)
        batch = tuple(t.to(dev, non_blocking=True) for t in batch)
        opt.zero_grad()
        out = model(batch)
        out["loss"].backward()
        if self.world > 1:
            from .parallel import DataParallel
            DataParallel(model, opt).sync_gradients()
        opt.step()
        return out

    # ------------------------------------------------------------------ fit (trainer.py:455-696)
    def fit(self, train_data, valid_data=None, saved=True, callback_fn=None):
        assert self.total_iters > 0, "config['total_iters'] must be set"
        self.model.train()
        iterator = iter(train_data)
        epoch_idx, total_loss, t0 = 0, 0.0, time.time()
        while self.train_step < self.total_iters:
            try:
                data = next(iterator)
            except StopIteration:                            # restart the generator for the next epoch (:498-505)
                epoch_idx += 1
                if hasattr(getattr(train_data, "sampler", None), "set_epoch"):
                    train_data.sampler.set_epoch(epoch_idx)
                iterator = iter(train_data)
                data = next(iterator)
            self.optimizer.set_lr(self.lr_at(self.train_step))   # LambdaLR: step k runs with lambda(k)
            out = self._step(data)
            loss = float(out["loss"].detach())
            if loss != loss:
                raise ValueError("Training loss is nan")     # trainer.py:371-373
            total_loss += loss
            self.train_step += 1
            if self.eval_interval > 0 and self.train_step % self.eval_interval == 0:
                self.train_loss_dict[self.train_step] = total_loss
                self._log(f"step {self.train_step} epoch {epoch_idx} train_loss {total_loss:.4f} "
                          f"lr {self.lr_at(self.train_step - 1):.3e} [{time.time() - t0:.1f}s]")
                total_loss = 0.0
                if not valid_data:
                    if saved:
                        self.save_checkpoint()
                    continue
                valid_result = self.evaluate(valid_data)
                valid_score = calculate_valid_score(valid_result, self.eval_pred_len, self.valid_metric)
                self.model.train()
                self.best_valid_score, self.no_improve_times, stop_flag, update_flag = early_stopping(
                    valid_score, self.best_valid_score, self.no_improve_times, max_step=self.stopping_step,
                    bigger=self.valid_metric_bigger)
                for p in self.metrics_pred_len_list:
                    row = dict(valid_result[f"pred_{p}"])
                    row["Pred Metrics"], row["Train Steps"] = f"pred_{p}", self.train_step
                    self.results.append(row)
                self._log(f"step {self.train_step} valid_score {valid_score:.6f} best {self.best_valid_score:.6f}")
                if update_flag:
                    if saved:
                        self.save_checkpoint()
                    self.best_valid_result = valid_result
                if callback_fn:
                    callback_fn(epoch_idx, valid_score)
                if stop_flag:
                    self._log(f"Finished training, best eval result at step "
                              f"{self.train_step - self.no_improve_times * self.eval_interval}")
                    break
        self._flush_step()
        self.model.eval()
        return self.best_valid_score, self.best_valid_result

    # ------------------------------------------------------------------ evaluate (trainer.py:698-729, 826-1153)
    def _flush_step(self):
        if self._stepper is not None and hasattr(self._stepper, "flush"):
            self._stepper.flush()                  # pending dense update of the sharded step

    @torch.no_grad()
    def evaluate(self, eval_data, all_item_tags=None, all_tags_NC=None):
        """Full-sort evaluation.  all_item_tags [C, N] (bool / 0-1; this rank's columns when the table is sharded)
        masks items per prior head; all_tags_NC [N, C] feeds the Entropy metric.  Returns
        {'pred_p': {metric: mean over users}} (+ 'shared')."""
        model, cfg = self.model, self.config
        self._flush_step()
        model.eval()
        dev = next(model.parameters()).device
        self.item_feature = model.compute_item_all()                       # trainer.py:790
        C = cfg["eval_num_cats"]
        N_local = self.item_feature.shape[0]
        if all_item_tags is None:
            all_item_tags = torch.ones(C, N_local, dtype=torch.bool, device=dev)    # batchset.py:36-38
        collector, evaluator = Collector(cfg), Evaluator(cfg)
        if all_tags_NC is not None:
            collector.set_all_tags(all_tags_NC)
        num_total = 0
        K = max(cfg["topk"])
        lockstep = self.sharded and self.world > 1       # sharded predict_topk is a collective: ranks stay in step
        it = iter(eval_data)
        while True:
            ev = next(it, None)
            if lockstep:
                more = torch.tensor([0 if ev is None else 1], dtype=torch.int32, device=dev)
                dist.all_reduce(more, op=dist.ReduceOp.MAX)
                if int(more.item()) == 0:
                    break
            elif ev is None:
                break
            if ev is None:
                # this rank ran out of users first: it still scores the other ranks' users against its rows
                Pe = cfg["eval_pred_len"]
                seq = torch.zeros((1, model.max_seq_length), dtype=torch.int64, device=dev)
                tt = torch.ones((1, Pe, C), dtype=torch.int64, device=dev)
                model.predict_topk(seq, self.item_feature, all_item_tags, tt, history_index=None, K=K)
                continue
            seq = ev["item_seq"].to(dev)
            tt = ev["target_tags"].to(dev)
            hu, hi = ev["history_index"]
            hist = (hu.to(dev), hi.to(dev)) if cfg.get("suppress_history", True) is not False else None
            top = model.predict_topk(seq, self.item_feature, all_item_tags, tt, history_index=hist, K=K)
            collector.eval_batch_collect(None, ev["positive_u"], ev["item_target"].to(dev), tt, topk=top)
            num_total += seq.shape[0]
        summary = {}
        if all_tags_NC is not None and evaluator.shared_metrics:
            summary["shared"] = self._reduce(evaluator.evaluate(collector.get_data_struct(-1), pred_len=-1), num_total)
        for p in self.metrics_pred_len_list:
            res = evaluator.evaluate(collector.get_data_struct(p), pred_len=p)
            summary[f"pred_{p}"] = self._reduce(res, num_total)
        return summary

    def _reduce(self, result, num_total):
        """Metric values are SUMS over this rank's users: all-reduce, then divide (trainer.py:1097-1123)."""
        out = {}
        dev = next(self.model.parameters()).device
        for k in sorted(result):
            v = result[k]
            is_tuple = isinstance(v, tuple)
            val, num = (v if is_tuple else (v, num_total))
            t = torch.tensor([float(val), float(num)], dtype=torch.float64, device=dev)
            if self.world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            out[k] = round(float(t[0]) / max(1.0, float(t[1])), self.metric_decimal_place)
        return out

    # ------------------------------------------------------------------ checkpoints (trainer.py:319-369)
    def _full_state_dict(self):
        """Reference-keyed state dict; a row-sharded table is gathered back to [N, D] (id g lives on rank g % W)."""
        sd = {k: v.detach().clone() for k, v in self.model.state_dict().items()}
        st = getattr(self.model, "sharded_table", None)
        if st is not None and st.W > 1:
            local = self.model.item_embedding.weight.data
            rows = st.local_rows(st.item_num, st.W, 0)
            pad = torch.zeros((rows, local.shape[1]), dtype=local.dtype, device=local.device)
            pad[:local.shape[0]] = local
            parts = [torch.empty_like(pad) for _ in range(st.W)]
            dist.all_gather(parts, pad, group=st.group)
            full = torch.empty((st.item_num, local.shape[1]), dtype=local.dtype, device=local.device)
            for r in range(st.W):
                n_r = st.local_rows(st.item_num, st.W, r)
                full[r::st.W] = parts[r][:n_r]
            sd["item_embedding.weight"] = full
        return sd

    def save_checkpoint(self, path=None):
        self._flush_step()
        sd = self._full_state_dict()
        opt = self.optimizer
        if self.rank != 0 and not self.sharded:
            return None
        path = path or os.path.join(self.checkpoint_dir or ".", f"b200rec-HSTU-rank{self.rank}.pth")
        names = {p: n for n, p in self.model.named_parameters()}
        opt_state = {names[p]: (m.detach().cpu(), v.detach().cpu()) for p, (m, v) in opt.state.items() if p in names}
        state = {"model": {k: v.cpu() for k, v in sd.items()},
                 "optimizer": {"state": opt_state, "step_count": opt.step_count, "shard": (self.rank, self.world)},
                 "config": dict(self.config), "iter_idx": self.train_step,
                 "best_valid_score": self.best_valid_score, "rng_state": torch.get_rng_state(),
                 # trainer.py:362-363 restores both generators; the Philox dropout stream is keyed by a step counter
                 "cuda_rng_state": torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
                 "dropout_step": None if self.model._rng_step is None else int(self.model._rng_step.item())}
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save(state, path)
        return path

    def load_checkpoint(self, path):
        state = torch.load(path, map_location="cpu", weights_only=False)
        model, opt = self.model, self.optimizer
        st = getattr(model, "sharded_table", None)
        sd = state["model"]
        if st is not None and st.W > 1:
            sd = dict(sd)
            sd["item_embedding.weight"] = st.shard_of(sd["item_embedding.weight"], st.W, st.rank)
        with torch.no_grad():
            for k, v in model.state_dict().items():
                v.copy_(sd[k].to(v.device))
        model.invalidate_shadows()
        names = dict(model.named_parameters())
        for n, (m, v) in state["optimizer"]["state"].items():
            if n in names and tuple(m.shape) == tuple(names[n].shape):
                pm, pv = opt._st(names[n])
                pm.copy_(m.to(pm.device))
                pv.copy_(v.to(pv.device))
        opt.step_count = state["optimizer"]["step_count"]
        if opt._coef is not None:
            opt._coef[3:4].fill_(float(opt.step_count))
        opt.mark_all_current()
        self.train_step = state["iter_idx"]
        self.best_valid_score = state["best_valid_score"]
        if state.get("rng_state") is not None:
            torch.set_rng_state(state["rng_state"])
        if state.get("cuda_rng_state") is not None and torch.cuda.is_available():
            torch.cuda.set_rng_state(state["cuda_rng_state"])
        if state.get("dropout_step") is not None:
            dev = next(model.parameters()).device
            if model._rng_step is None or model._rng_step.device != dev:
                model._rng_step = torch.zeros(1, dtype=torch.int64, device=dev)
            model._rng_step.fill_(int(state["dropout_step"]))      # in place: captured graphs read this buffer
        return state                               # captured graphs read parameters / state in place: still valid
