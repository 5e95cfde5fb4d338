"""src.utils.utils_eval — B200 drop-in (reference: src/utils/utils_eval.py)."""
from cddpm.eval_tail import (_test_end, _test_step, apply_3d_median_filter, apply_brainmask_volume,  # noqa: F401
                             compute_prc, compute_roc, dice, filter_3d_connected_components, find_best_val, fpr,
                             get_eval_dictionary, tpr)
