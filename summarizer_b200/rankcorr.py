"""Rank correlation (utils/eval.py:49-72) — device implementation."""


def rank_correlation(machine_scores, user_scores, metric="spearmanr"):
    raise NotImplementedError("evaluate_scores: the sm_100a rank-correlation kernel is not built yet "
                              "(SURVEY.md §8f NEXT-1); there is no CPU fallback")
