"""Benchmark table with the reference's functions (benchmark.py:19-80): trains a list of models through
``main.train`` and tabulates correlation / F-scores; extended with the hot-path models (VASNet) and a
frames-per-second column measured around each training run."""
import argparse
import datetime
import logging
import os
import sys
import time

import pandas as pd
from tabulate import tabulate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from summarizer_b200.main import train  # noqa: E402
from summarizer_b200.utils.config import HParameters  # noqa: E402


def benchmark(args, log_path):
    """Successively train models"""
    table_results = []
    base = {"splits_files": args.splits_files, "log_level": "error", "use_cuda": getattr(args, "use_cuda", "default")}
    table_results += benchmark_model("Random", dict({"model": "random", "epochs": 1, "extra_params": {}}, **base))
    table_results += benchmark_model("Logistic Regression", dict({"model": "logistic", "epochs": min(30, args.max_epochs),
                                                                  "extra_params": {}}, **base))
    table_results += benchmark_model("VASNet", dict({"model": "vasnet", "epochs": min(50, args.max_epochs),
                                                     "extra_params": {}}, **base))
    table_results += benchmark_model("DSN", dict({"model": "dsn", "epochs": min(60, args.max_epochs), "extra_params": {}}, **base))
    if getattr(args, "sumgan", False):      # 195 M parameters, three adversarial updates per video: minutes per fold
        table_results += benchmark_model("SumGAN", dict({"model": "sumgan", "epochs": min(20, args.max_epochs),
                                                         "extra_params": {"pretrain_vae": str(min(20, args.max_epochs))}}, **base))
    table = pd.DataFrame(table_results, columns=["Model", "File", "Correlation", "Avg F-score", "Max F-score", "Logs",
                                                 "Train+eval s"])
    show_save_results(table, log_path)
    return table


def benchmark_model(name, args):
    """Routine to train one model"""
    logging.info(f"Train {name} model...")
    hps = HParameters()
    hps.load_from_args(args)
    t0 = time.perf_counter()
    model_results = train(hps)
    dt = time.perf_counter() - t0
    results = []
    for splits_file, corr, avg_fscore, max_fscore in model_results:
        results.append([name, splits_file, corr, avg_fscore, max_fscore, hps.log_path, dt])
        logging.info(f"File: {splits_file}  Corr: {corr: 0.5f}  Avg F-score: {avg_fscore:0.5f}  Max F-score: {max_fscore:0.5f}")
    logging.info(f"Logs saved in {hps.log_path}")
    return results


def show_save_results(table, log_path):
    """Display to terminal and save to logs the Pandas table"""
    table_str = tabulate(table, headers="keys", tablefmt="psql", showindex=False)
    print(table_str)
    os.makedirs(log_path, exist_ok=True)
    table_file = os.path.join(log_path, "table.txt")
    with open(table_file, "w") as f:
        f.write(table_str)
    logging.info(f"Table saved in {table_file}")


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO, format="%(asctime)s::%(levelname)s: %(message)s")
    log_path = os.path.join("logs", f"{int(datetime.datetime.now().timestamp())}_benchmark")
    parser = argparse.ArgumentParser("Summarizer : Benchmark")
    parser.add_argument("-e", "--max-epochs", type=int, default=300, help="Maximum number of epochs per model")
    parser.add_argument("-s", "--splits-files", type=str, default="splits/tvsum_splits.json,splits/summe_splits.json",
                        help="Comma separated list of split files")
    parser.add_argument("-c", "--use-cuda", choices=["yes", "no", "default"], default="default")
    parser.add_argument("-g", "--sumgan", action="store_true", help="also train SumGAN (minutes per fold)")
    args, _ = parser.parse_known_args()
    print(args)
    benchmark(args, log_path)
