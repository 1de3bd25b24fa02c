"""Split-driven training / cross-validation driver with the reference's ``train(hps)`` contract and CLI
(main.py:10-104): for every split file, for every fold: reset -> train -> keep the best fold's weights; then
cross-validation means, TensorBoard hparams, predictions over the whole dataset.

Multi-GPU (one process per GPU, ``torchrun --nproc-per-node N main.py ...``): the (split file, fold) jobs are
independent (main.py:14,26), so ALL of them — across split files — are dealt to the ranks longest-first and run
with NO data-path collective; rank 0 gathers the per-fold results (and the best fold's weights) through
torch.distributed and returns the same ``results`` list as a single-process run."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from summarizer_b200.utils import Proportion  # noqa: E402
from summarizer_b200.utils.config import HParameters  # noqa: E402


def plan_folds(costs, world):
    """Longest-processing-time assignment of jobs to ranks.  ``costs``: list of (job, cost).
    Returns one job list per rank; deterministic (ties by job order)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i][1], i))
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(costs[i][0])
        load[r] += costs[i][1]
    return [sorted(jobs) for jobs in out]


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist, dist.get_rank(), dist.get_world_size()
    return None, 0, 1


def fold_cost(hps, model, splits_file, fold):
    """epochs x number of training steps (frames) of the fold — the scheduling weight."""
    keys = hps.splits_of_file[splits_file][fold]["train_keys"]
    return float(hps.epochs) * sum(int(np.shape(model.dataset[k]["features"])[0]) for k in keys)


def _cpu_state(state_dict):
    """Detached host copy of a state dict (``state_dict()`` aliases the live parameters, which the next fold resets)."""
    return {k: v.detach().to("cpu", copy=True) for k, v in state_dict.items()}


def _train_jobs_concurrently(hps, models, jobs, k, rank, world):
    """Fold-concurrent training on ONE GPU (``--concurrent_folds K``): a batch-1 training step keeps 12-24
    of the 148 SMs busy per GEMM, and the folds of a cross-validation are independent models (main.py:26), so K of this
    rank's (split file, fold) jobs train side by side — K worker threads, each with its own trainer (model, optimizer,
    per-video step graphs, resident dataset) and its own CUDA stream, pulling jobs longest-first from one queue.  No
    result depends on the interleaving: every fold sees exactly the reference's loop.  Returns (fold_results, best) as
    the sequential loop builds them (best fold = highest correlation, first fold on ties, with a host copy of its
    weights)."""
    import queue
    import threading
    costs = {(i, f): fold_cost(hps, models[hps.splits_files[i]], hps.splits_files[i], f) for i, f in jobs}
    todo = queue.SimpleQueue()
    for job in sorted(jobs, key=lambda j: (-costs[j], j)):
        todo.put(job)
    import contextlib
    on_gpu = bool(hps.use_cuda) and torch.cuda.is_available()     # host-only models (Rand / Logistic on the CPU): plain threads
    device = torch.cuda.current_device() if on_gpu else None
    lock = threading.Lock()
    fold_results, best, errors = {}, {}, []

    def worker(w):
        try:
            stream = None
            if on_gpu:
                torch.cuda.set_device(device)
                stream = torch.cuda.Stream(device=device)
            mine = dict(models) if w == 0 else {}                 # worker 0 reuses the trainers the caller built
            with (torch.cuda.stream(stream) if on_gpu else contextlib.nullcontext()):
                while not errors:
                    try:
                        i, fold = todo.get_nowait()
                    except queue.Empty:
                        break
                    sf = hps.splits_files[i]
                    if sf not in mine:
                        mine[sf] = hps.model_class(hps, sf)
                    model = mine[sf]
                    corr, avg_f, max_f = model.reset().train(fold)
                    if model.best_weights is None:
                        raise Exception("best_weights property is empty, can't save model's weights")
                    state = _cpu_state(model.best_weights)
                    with lock:
                        fold_results[(i, fold)] = (float(corr), float(avg_f), float(max_f))
                        cur = best.get(i)
                        if cur is None or (-float(corr), fold) < (-cur[0], cur[1]):
                            best[i] = (float(corr), fold, state)
                    n_folds = len(hps.splits_of_file[sf])
                    hps.logger.info(f"File: {sf}   Fold: {fold+1}/{n_folds}   Corr: {corr: 0.5f}  "
                                    f"Avg F-score: {avg_f:0.5f}  Max F-score: {max_f:0.5f}" + (f" (rank {rank})" if world > 1 else ""))
                if stream is not None:
                    stream.synchronize()
        except BaseException as e:                                 # surfaced by the caller: no silent loss of a fold
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(w,), name=f"smz-fold-{w}") for w in range(k)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return fold_results, best


def train(hps):
    """Training.  Returns [(splits_file, mean corr, mean avg F, mean max F), ...] (main.py:10-72).

    One process: the reference's loop.  Several ranks (fold-parallel): ALL (split file, fold) jobs of the run are
    dealt to the ranks at once, longest first (SumMe + TVSum = 10 jobs, the TVSum folds ~3x heavier), every rank
    trains its jobs with no collective, then the per-fold results are gathered and, per split file, the rank that
    owns the best fold (highest correlation, first fold on ties as main.py:33-35) ships that fold's weights to
    rank 0 through the process group — no files, no dependence on a shared log directory."""
    dist, rank, world = _dist()
    data_parallel = bool((hps.extra_params or {}).get("data_parallel", False))
    if data_parallel:
        world = 1            # every rank trains every fold together (gradient all-reduce inside the trainer)
    models = {sf: hps.model_class(hps, sf) for sf in hps.splits_files}
    jobs = [(i, f) for i, sf in enumerate(hps.splits_files) for f in range(len(hps.splits_of_file[sf]))]
    if world > 1:
        costs = [((i, f), fold_cost(hps, models[hps.splits_files[i]], hps.splits_files[i], f)) for i, f in jobs]
        mine = plan_folds(costs, world)[rank]
    else:
        mine = jobs

    fold_results = {}                      # (file index, fold) -> (corr, avg F, max F)
    best = {}                              # file index -> (corr, fold, host copy of the weights)
    concurrent = int((hps.extra_params or {}).get("concurrent_folds", 1))
    if concurrent > 1 and not data_parallel and len(mine) > 1:
        fold_results, best = _train_jobs_concurrently(hps, models, mine, min(concurrent, len(mine)), rank, world)
        if world == 1:
            for i, (_, _, state) in best.items():
                torch.save(state, hps.weights_path[hps.splits_files[i]])
        mine = []
    for i, fold in mine:
        sf = hps.splits_files[i]
        n_folds = len(hps.splits_of_file[sf])
        if fold == 0 or world > 1:
            hps.logger.info(f"Start training on {sf}" + (f" (rank {rank})" if world > 1 else ""))
        model = models[sf]
        best_corr, best_avg_f, best_max_f = model.reset().train(fold)
        fold_results[(i, fold)] = (float(best_corr), float(best_avg_f), float(best_max_f))
        if best_corr > best.get(i, (-1.0,))[0]:
            if world > 1:
                if model.best_weights is None:
                    raise Exception("best_weights property is empty, can't save model's weights")
                best[i] = (float(best_corr), fold, _cpu_state(model.best_weights))
            else:
                best[i] = (float(best_corr), fold, None)
                if rank == 0:
                    model.save_best_weights(hps.weights_path[sf])
        hps.logger.info(f"File: {sf}   Fold: {fold+1}/{n_folds}   Corr: {best_corr: 0.5f}  "
                        f"Avg F-score: {best_avg_f:0.5f}  Max F-score: {best_max_f:0.5f}")

    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (fold_results, {i: b[:2] for i, b in best.items()}))
        fold_results = {k: r for g, _ in gathered for k, r in g.items()}
        for i, sf in enumerate(hps.splits_files):
            cand = [(-b[i][0], b[i][1], r) for r, (_, b) in enumerate(gathered) if i in b]
            if not cand:
                raise RuntimeError(f"{sf}: no rank reported a trained fold")
            _, _, owner = min(cand)                               # highest correlation, then the first fold
            box = [best[i][2] if rank == owner else None]
            if owner != 0:                                        # owner -> rank 0
                dist.broadcast_object_list(box, src=owner)
            if rank == 0:
                if box[0] is None:
                    raise RuntimeError(f"{sf}: the best fold's weights did not arrive from rank {owner}")
                torch.save(box[0], hps.weights_path[sf])

    results = []
    for i, sf in enumerate(hps.splits_files):
        n_folds = len(hps.splits_of_file[sf])
        weights_path, pred_path = hps.weights_path[sf], hps.pred_path[sf]
        missing = [f for f in range(n_folds) if (i, f) not in fold_results]
        if missing:
            raise RuntimeError(f"{sf}: folds {missing} were not trained by any rank")
        corrs_cv = [fold_results[(i, f)][0] for f in range(n_folds)]
        avg_fscores_cv = [fold_results[(i, f)][1] for f in range(n_folds)]
        max_fscores_cv = [fold_results[(i, f)][2] for f in range(n_folds)]
        hps.logger.info(f"File: {sf}   Cross-validation Corr: {np.mean(corrs_cv): 0.5f}  "
                        f"Avg F-score: {np.mean(avg_fscores_cv):0.5f}  Max F-score: {np.mean(max_fscores_cv):0.5f}")
        if rank == 0:
            if not os.path.exists(weights_path):
                raise FileNotFoundError(f"{sf}: best weights {weights_path} were not written")
            hps.logger.info(f"File: {sf}   Best weights: {weights_path}")
            hparam_dict = hps.get_full_hps_dict()
            hparam_dict["dataset"] = hps.dataset_name_of_file[sf]
            metric_dict = {f"F-score_max/Fold_{f+1}": s for f, s in enumerate(max_fscores_cv)}   # main.py:56-58 keeps the last
            metric_dict["Correlation/CV_Average"] = np.mean(corrs_cv)
            metric_dict["F-score_avg/CV_Average"] = np.mean(avg_fscores_cv)
            metric_dict["F-score_max/CV_Average"] = np.mean(max_fscores_cv)
            hps.writer.add_hparams(hparam_dict, metric_dict)
            model = models[sf]
            model.reset().load_weights(weights_path)
            model.best_weights = model.model.state_dict()
            model.predict_dataset(pred_path)
            hps.logger.info(f"File: {sf}   Machine predictions: {pred_path}")
        results.append((sf, np.mean(corrs_cv), np.mean(avg_fscores_cv), np.mean(max_fscores_cv)))
    return results


def parse_extra(unknown_args):
    """main.py:91 — unknown ``--key value`` pairs / ``--flag`` become extra_params strings / True."""
    if not unknown_args:
        return {}
    nxt = unknown_args[1:] + ["-"]
    return {unknown_args[i].lstrip("-"): (u.lstrip("-") if u[0] != "-" else True)
            for i, u in enumerate(nxt) if unknown_args[i][0] == "-"}


def build_parser():
    parser = argparse.ArgumentParser("Summarizer : Model Training")
    parser.add_argument("-c", "--use-cuda", choices=["yes", "no", "default"], default="default", help="Use cuda for pytorch models")
    parser.add_argument("-i", "--cuda-device", type=int, help="If cuda-enabled, ID of GPU to use")
    parser.add_argument("-s", "--splits-files", type=str, help="Comma separated list of split files (shorthands: minimal, overfit, all)")
    parser.add_argument("-m", "--model", type=str, help="Model class name")
    parser.add_argument("-e", "--epochs", type=int, help="Number of epochs for train mode")
    parser.add_argument("-r", "--lr", type=float, help="Learning rate for train mode")
    parser.add_argument("-d", "--weight-decay", type=float, help="Weight decay (L2 penalty-based regularization)")
    parser.add_argument("-t", "--test-every-epochs", type=int, help="Evaluate the model every nth epoch on the current fold's validation set")
    parser.add_argument("-p", "--summary-proportion", type=float, choices=Proportion(), help="Length of video summary (as a proportion of original video length)")
    parser.add_argument("-a", "--selection-algorithm", choices=["knapsack", "rank"], help="Keyshot selection algorithm to build the summary video")
    parser.add_argument("-l", "--log-level", choices=["critical", "error", "warning", "info", "debug"], default="info", help="Set logger to custom level")
    return parser


def main(argv=None):
    args, unknown_args = build_parser().parse_known_args(argv)
    hps_init = dict(args.__dict__)
    hps_init["extra_params"] = parse_extra(unknown_args)
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", 0))
        # fold-parallel runs only gather small Python objects (and one state dict): a host-side gloo group is enough
        # and spares the NCCL communicator set-up; --data_parallel all-reduces gradients and needs NCCL
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
            hps_init["cuda_device"] = local
        if torch.cuda.is_available() and hps_init["extra_params"].get("data_parallel", False):
            if str(hps_init["extra_params"].get("dp_cuda_graphs", "no")).lower() in ("yes", "1", "true"):
                # replicas capture their step graphs at DIFFERENT steps: NCCL must not turn a capture into a collective
                # host-side operation (graph-time buffer registration exchanges handles between the ranks; for NVLS it is
                # collective).  A precaution — the 8-GPU run of this mode is still to be confirmed.
                os.environ.setdefault("NCCL_GRAPH_REGISTER", "0")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group("gloo")
    hps = HParameters()
    hps.load_from_args(hps_init)
    print("Hyperparameters:")
    print("----------------------------------------------------------------------")
    print(hps)
    print("----------------------------------------------------------------------")
    results = train(hps)
    hps.writer.close()
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    return results


if __name__ == "__main__":
    main()
