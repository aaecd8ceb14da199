"""Host-side box ordering / word segmentation / transcription post-processing and the CER metric.

These stay on the host in the reference too: they are
O(chars per line) Python on a few dozen boxes.  Semantics follow the reference exactly; tests pin
them against the live reference functions (tests/test_oracle.py::test_textproc_matches_golden, tests/golden/textproc_golden.json).
"""
from __future__ import annotations

import string

LARGE_NUMBER = 1_000_000  # infer_effocr.py:238 `LARGE_NUM`, infer_effocr_onnx_multi.py `LARGE_NUMBER`


def flatten(x):
    """utils/spell_check_utils.py:68-73: flatten one level of tuples/lists, strings stay atoms."""
    for e in x:
        if isinstance(e, (list, tuple)):
            yield from flatten(e)
        else:
            yield e


def create_distinct_lowercase():
    """utils/spell_check_utils.py:60-61."""
    return list("aenr")


def create_nondistinct_lowercase():
    """utils/spell_check_utils.py:64-65."""
    return list("wuosvcxz")


def _key(vertical):
    return (lambda b: b[1]) if vertical else (lambda b: b[0])


def en_preprocess(bboxes_char, bboxes_word, vertical: bool = False, score_thresh=None, score_thresh_word=None):
    """Sort char / word boxes along the text direction and find, for every word box, the index of the
    char whose right edge is nearest to (and right of) the word's left edge.

    ONNX path (infer_effocr_onnx_multi.py:70-89): no score filtering (score_thresh=None).
    Torch path (infer_effocr.py:346-368): boxes carry a 5th score column, filtered with `>` AFTER
    sorting, and only x0..y1 are kept.
    """
    sorted_char = sorted(bboxes_char, key=_key(vertical))
    sorted_word = sorted(bboxes_word, key=_key(vertical))
    if score_thresh is not None:
        sorted_char = [b[:4] for b in sorted_char if b[4] > score_thresh]
    if score_thresh_word is not None:
        sorted_word = [b[:4] for b in sorted_word if b[4] > score_thresh_word]
    word_end_idx = []
    closest_idx = 0
    char_rights = [b[2] for b in sorted_char]
    word_lefts = [b[0] for b in sorted_word]
    for wordleft in word_lefts:
        prev_dist = LARGE_NUMBER
        for idx, charright in enumerate(char_rights):
            dist = abs(wordleft - charright)
            if dist < prev_dist and charright > wordleft:
                prev_dist = dist
                closest_idx = idx
        word_end_idx.append(closest_idx)  # note: closest_idx carries over between words, as in the reference
    assert len(word_end_idx) == len(sorted_word)
    return sorted_char, word_end_idx


def en_preprocess_np(char_boxes, word_boxes, vertical: bool = False):
    """`en_preprocess` of the ONNX path (no score filtering) on float32 arrays [n, >=4] / [w, >=4], vectorised: the same
    stable ordering and, per word, the same float32 `abs(wordleft - charright)` values, strict `<` (first minimum wins),
    `charright > wordleft` filter, 1e6 starting distance and carry-over of the previous word's index as the scalar
    loop above -- one [w, n] array operation instead of w * n Python iterations on numpy scalars.
    -> (char boxes sorted along the text direction [n, :], word_end_idx list)."""
    import numpy as np

    char_boxes = np.asarray(char_boxes, dtype=np.float32)
    word_boxes = np.asarray(word_boxes, dtype=np.float32)
    if char_boxes.ndim != 2:
        char_boxes = char_boxes.reshape(-1, 4)
    if word_boxes.ndim != 2:
        word_boxes = word_boxes.reshape(-1, 4)
    k = 1 if vertical else 0
    sorted_char = char_boxes[np.argsort(char_boxes[:, k], kind="stable")] if len(char_boxes) else char_boxes
    if len(word_boxes) == 0:
        return sorted_char, []
    sorted_word = word_boxes[np.argsort(word_boxes[:, k], kind="stable")]
    word_end_idx, closest_idx = [], 0
    if len(sorted_char):
        rights, lefts = sorted_char[:, 2], sorted_word[:, 0]
        dist = np.abs(lefts[:, None] - rights[None, :])  # float32, as the scalar subtraction
        ok = (rights[None, :] > lefts[:, None]) & (dist < np.float32(LARGE_NUMBER))
        best = np.where(ok, dist, np.float32(np.inf)).argmin(axis=1)  # first index of the minimum
        found = ok.any(axis=1)
        for j in range(len(sorted_word)):
            if found[j]:
                closest_idx = int(best[j])
            word_end_idx.append(closest_idx)
    else:
        word_end_idx = [0] * len(sorted_word)
    return sorted_char, word_end_idx


def jp_preprocess(bboxes_char, vertical: bool = True, score_thresh=None):
    """infer_effocr_onnx_multi.py:134-140 / infer_effocr.py:413-419."""
    sorted_char = sorted(bboxes_char, key=_key(vertical))
    if score_thresh is not None:
        sorted_char = [b[:4] for b in sorted_char if b[4] > score_thresh]
    return sorted_char


def en_postprocess(line_output, word_end_idx, charheights, charbottoms, anchor_margin=None, anchor_multiplier=4,
                   spell_checker=None):
    """infer_effocr.py:371-410 / infer_effocr_onnx_multi.py:92-131: prefix a space to the chars at
    word_end_idx; optional height/bottom based case and period fixes when anchor_margin is set."""
    assert len(line_output) == len(charheights) == len(charbottoms), \
        f"{len(line_output)} == {len(charheights)} == {len(charbottoms)}; {line_output}; {charbottoms}; {charheights}"
    if any(map(lambda x: len(x) == 0, (line_output, word_end_idx, charheights, charbottoms))):
        return None
    ends = set(word_end_idx)
    outchars_w_spaces = [" " + x if idx in ends else x for idx, x in enumerate(line_output)]
    if anchor_margin is None and spell_checker is None and all(isinstance(x, str) and len(x) == 1 and not x.isspace() for x in line_output):
        # fast path of the ONNX driver's call (single non-blank characters, no height rules): the two *_w_spaces lists are
        # only length-checked below, and their length is n + (inserted markers) - (a leading marker, dropped)
        # (the reference's length assertion can only trip when the first height equals the marker value itself)
        assert 0 in ends or charheights[0] != LARGE_NUMBER, f"charheights = {charheights}; output = {line_output}"
        return "".join(outchars_w_spaces).strip()
    charheights_w_spaces, charbottoms_w_spaces = [], []
    for idx, (hgt, bot) in enumerate(zip(charheights, charbottoms)):
        if idx in ends:
            charheights_w_spaces.append(LARGE_NUMBER)
            charbottoms_w_spaces.append(0)
        charheights_w_spaces += list(flatten([hgt])) if isinstance(hgt, (list, tuple)) else [hgt]
        charbottoms_w_spaces += list(flatten([bot])) if isinstance(bot, (list, tuple)) else [bot]
    charbottoms_w_spaces = charbottoms_w_spaces[1:] if charbottoms_w_spaces[0] == 0 else charbottoms_w_spaces
    charheights_w_spaces = charheights_w_spaces[1:] if charheights_w_spaces[0] == LARGE_NUMBER else charheights_w_spaces
    line_output = "".join(outchars_w_spaces).strip()
    assert len(charheights_w_spaces) == len(line_output), \
        f"charheights_w_spaces = {len(charheights_w_spaces)}; output = {len(line_output)}; {charheights_w_spaces}; {line_output}"
    distinct = create_distinct_lowercase()
    output_distinct_lower_idx = [idx for idx, c in enumerate(line_output) if c in distinct]
    fix = len(output_distinct_lower_idx) > 0 and anchor_margin is not None
    if fix:
        avg_h = sum(charheights_w_spaces[idx] for idx in output_distinct_lower_idx) / len(output_distinct_lower_idx)
        tolower = [idx for idx, c in enumerate(line_output) if abs(charheights_w_spaces[idx] - avg_h) < anchor_margin * avg_h]
        toupper = [idx for idx, c in enumerate(line_output)
                   if charheights_w_spaces[idx] - avg_h > anchor_margin * anchor_multiplier * avg_h]
        avg_b = sum(charbottoms_w_spaces[idx] for idx in output_distinct_lower_idx) / len(output_distinct_lower_idx)
        toperiod = [idx for idx, c in enumerate(line_output)
                    if c == "-" and abs(charbottoms_w_spaces[idx] - avg_b) < anchor_margin * avg_h]
    if spell_checker is not None:
        line_output = spell_checker(line_output)
    if fix:
        nondistinct = create_nondistinct_lowercase()
        line_output = "".join([c.lower() if idx in tolower else c for idx, c in enumerate(line_output)])
        line_output = "".join([c.upper() if idx in toupper and c in nondistinct else c for idx, c in enumerate(line_output)])
        line_output = "".join(["." if idx in toperiod else c for idx, c in enumerate(line_output)])
    return line_output


# ------------------------------------------------------------------ metric (utils/eval_utils.py)
def edit_distance(a: str, b: str) -> int:
    """Levenshtein distance (nltk.metrics.distance.edit_distance default: unit costs, no transposition)."""
    if a == b:
        return 0
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def string_cleaner(s: str) -> str:
    """utils/eval_utils.py:14-22."""
    return (s.replace("“", "\"").replace("”", "\"").replace("''", "\"").replace("‘‘", "\"").replace("’’", "\"")
            .replace("\n", ""))


def textline_evaluation(pairs, print_incorrect=False, no_spaces_in_eval=False, norm_edit_distance=False, uncased=False):
    """utils/eval_utils.py:25-70 -> (textline accuracy in percent, CER); pairs are (ground truth, prediction)."""
    n_correct = 0
    edit_count = 0
    length_of_data = len(pairs)
    n_chars = sum(len(gt) for gt, _ in pairs)
    for gt, pred in pairs:
        pred, gt = string_cleaner(pred), string_cleaner(gt)
        gt = gt.strip() if not no_spaces_in_eval else gt.strip().replace(" ", "")
        pred = pred.strip() if not no_spaces_in_eval else pred.strip().replace(" ", "")
        if uncased:
            pred, gt = pred.lower(), gt.lower()
        if pred == gt:
            n_correct += 1
        elif print_incorrect:
            print(f"GT: {gt}\nPR: {pred}\n")
        if norm_edit_distance:  # ICDAR2019 normalised edit distance
            if len(gt) > len(pred):
                edit_count += edit_distance(pred, gt) / len(gt)
            else:
                edit_count += edit_distance(pred, gt) / len(pred)
        else:
            edit_count += edit_distance(pred, gt)
    accuracy = n_correct / float(length_of_data) * 100
    if norm_edit_distance:
        cer = edit_count / float(length_of_data)
    else:
        cer = edit_count / n_chars
    return accuracy, cer
