from jamie_b200.evaluation import (foscttm, label_transfer_accuracy, imputation_correlation, mean_feature_r,  # noqa: F401
                                   test_closer, test_LabelTA)
