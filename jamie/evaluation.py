from jamie_b200.evaluation import foscttm, label_transfer_accuracy, imputation_correlation  # noqa: F401
