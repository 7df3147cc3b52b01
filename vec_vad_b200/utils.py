"""Host-side helpers of the reference's ``utils.py`` and of the scoring stage of ``test.py`` (integer / index paths that
must stay bit-exact; metrics are plain sklearn like the reference)."""
import numpy as np


def calc_block_idx(x_min, x_max, y_min, y_max, h_step, w_step, mode):
    """Spatial blocks a bbox belongs to (utils.py:5-26): its centre (mode 1), plus the four edge mid-points (mode > 1),
    plus the four corners (mode >= 9), each averaged with the centre, divided by the block size and truncated to int."""
    cy, cx = (y_min + y_max) / 2, (x_min + x_max) / 2
    pts = [(cy, cx)]
    if mode > 1:
        pts += [(y_min, cx), (y_max, cx), (cy, x_min), (cy, x_max)]
    if mode >= 9:
        pts += [(y_min, x_min), (y_max, x_max), (y_max, x_min), (y_min, x_max)]
    arr = (np.array(pts, dtype=np.float64) + np.array([cy, cx], dtype=np.float64)) / 2
    hb = (arr[:, 0] / h_step).astype(int)
    wb = (arr[:, 1] / w_step).astype(int)
    return list(set(zip(list(hb), list(wb))))


def paint_score_mask(pixel_results, scores, bboxes, big_number=100000):
    """Running max of per-bbox score rectangles into a frame-sized map (test.py:350-357): ceil on all four bbox edges."""
    for m in range(len(scores)):
        box = bboxes[m]
        x_min, x_max = int(np.ceil(box[0])), int(np.ceil(box[2]))
        y_min, y_max = int(np.ceil(box[1])), int(np.ceil(box[3]))
        region = pixel_results[y_min:y_max, x_min:x_max]
        np.maximum(region, scores[m], out=region)
    return pixel_results


def save_roc_pr_curve_data(scores, labels, file_path, verbose=True):
    """Frame-level ROC / PR curves, AUROC and EER written to ``file_path`` (.npz) -- utils.py:29-65; returns the AUROC."""
    from sklearn.metrics import auc, precision_recall_curve, roc_curve
    scores, labels = scores.flatten(), labels.flatten()
    scores_pos, scores_neg = scores[labels == 1], scores[labels != 1]
    truth = np.concatenate((np.zeros_like(scores_neg), np.ones_like(scores_pos)))
    preds = np.concatenate((scores_neg, scores_pos))
    fpr, tpr, roc_thresholds = roc_curve(truth, preds)
    roc_auc = auc(fpr, tpr)
    # equal error rate: where the false-negative rate meets the false-positive rate
    fnr = 1 - tpr
    eer1 = fpr[np.nanargmin(np.absolute(fnr - fpr))]
    eer2 = fnr[np.nanargmin(np.absolute(fnr - fpr))]
    precision_norm, recall_norm, pr_thresholds_norm = precision_recall_curve(truth, preds)
    pr_auc_norm = auc(recall_norm, precision_norm)
    precision_anom, recall_anom, pr_thresholds_anom = precision_recall_curve(truth, -preds, pos_label=0)
    pr_auc_anom = auc(recall_anom, precision_anom)
    if verbose:
        print('AUC@ROC is {}'.format(roc_auc), 'EER1 is {}'.format(eer1), 'EER2 is {}'.format(eer2))
    np.savez_compressed(file_path, preds=preds, truth=truth, fpr=fpr, tpr=tpr, roc_thresholds=roc_thresholds, roc_auc=roc_auc,
                        precision_norm=precision_norm, recall_norm=recall_norm, pr_thresholds_norm=pr_thresholds_norm,
                        pr_auc_norm=pr_auc_norm, precision_anom=precision_anom, recall_anom=recall_anom,
                        pr_thresholds_anom=pr_thresholds_anom, pr_auc_anom=pr_auc_anom)
    return roc_auc
