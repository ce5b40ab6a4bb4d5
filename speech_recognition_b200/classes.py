"""Class lists, orders and label maps of the reference (classes.py:5-41,
input_data.py:49-60, make_submission.py:16-31, convert_from_see_v3_bugfix.py:67)."""
from collections import OrderedDict

import numpy as np

SILENCE_LABEL = '_silence_'
UNKNOWN_WORD_LABEL = '_unknown_'
AUDIO_NAMES = ['silence', 'unknown', 'yes', 'no', 'up', 'down', 'left', 'right', 'on', 'off', 'stop', 'go']


def prepare_words_list(wanted_words):
    return [SILENCE_LABEL, UNKNOWN_WORD_LABEL] + list(wanted_words)


def get_classes(wanted_only=False, extend_reversed=False):
    if wanted_only:
        classes = 'stop down off right up go on yes left no'.split(' ')
    else:
        classes = ('sheila nine stop bed four six down bird marvin cat off right seven eight up three '
                   'happy go zero on wow dog yes five one tree house two left no').split(' ')
    if extend_reversed:
        assert not wanted_only
        classes.extend(['new_owt', 'new_yppah', 'new_xis', 'new_esuoh', 'new_neves', 'new_thgie',
                        'new_ruof', 'new_tac', 'new_nivram', 'new_enin', 'new_aliehs', 'new_eert',
                        'new_orez', 'new_eerht', 'new_evif', 'new_deb', 'new_drib'])
    return classes


def get_int2label(wanted_only=False, extend_reversed=False):
    classes = prepare_words_list(get_classes(wanted_only, extend_reversed))
    return OrderedDict((i, l) for i, l in enumerate(classes))


def get_label2int(wanted_only=False, extend_reversed=False):
    classes = prepare_words_list(get_classes(wanted_only, extend_reversed))
    return OrderedDict((l, i) for i, l in enumerate(classes))


def map_to_valid(labels):
    return ['silence' if l == SILENCE_LABEL else 'unknown' if l == UNKNOWN_WORD_LABEL else l for l in labels]


def map_to_wanted(labels, wanted_words):
    return [l if l in wanted_words or l == 'silence' else 'unknown' for l in labels]


def class_map_32_to_12(order='heng'):
    """Destination column of each of the 32 classes in the 12-class vector.
    'heng'   -> AUDIO_NAMES order (convert_from_see_v3_bugfix.py:67-92)
    'frozen' -> silence, unknown, wanted words in training order (freeze_graph_32_classes.py:55-69)."""
    names32 = prepare_words_list(get_classes(False))
    if order == 'heng':
        target = AUDIO_NAMES
    elif order == 'frozen':
        wanted = get_classes(True)
        target = ['silence', 'unknown'] + [w for w in get_classes(False) if w in wanted]
    else:
        raise ValueError(order)
    cmap = np.ones(len(names32), np.int32)          # default: the 'unknown' group
    for i, nm in enumerate(names32):
        if nm == SILENCE_LABEL:
            cmap[i] = 0
        elif nm in target:
            cmap[i] = target.index(nm)
    return cmap
