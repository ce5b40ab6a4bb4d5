"""``prepare_model_settings`` -- same keys and arithmetic as the reference
(model.py:1785-1829)."""


def prepare_model_settings(label_count, sample_rate, clip_duration_ms, window_size_ms, window_stride_ms,
                           dct_coefficient_count, num_log_mel_features, output_representation='raw'):
    desired_samples = int(sample_rate * clip_duration_ms / 1000)
    window_size_samples = int(sample_rate * window_size_ms / 1000)
    window_stride_samples = int(sample_rate * window_stride_ms / 1000)
    length_minus_window = desired_samples - window_size_samples
    spectrogram_frequencies = 257
    if length_minus_window < 0:
        spectrogram_length = 0
    else:
        spectrogram_length = 1 + int(length_minus_window / window_stride_samples)
    if output_representation in ('mfcc', 'mfcc_and_raw'):
        fingerprint_size = num_log_mel_features * spectrogram_length
    elif output_representation == 'raw':
        fingerprint_size = desired_samples
    elif output_representation == 'spec':
        fingerprint_size = spectrogram_frequencies * spectrogram_length
    else:
        raise ValueError("Invalid output_representation: %s" % output_representation)
    return {
        'desired_samples': desired_samples,
        'window_size_samples': window_size_samples,
        'window_stride_samples': window_stride_samples,
        'spectrogram_length': spectrogram_length,
        'spectrogram_frequencies': spectrogram_frequencies,
        'dct_coefficient_count': dct_coefficient_count,
        'fingerprint_size': fingerprint_size,
        'label_count': label_count,
        'sample_rate': sample_rate,
        'num_log_mel_features': num_log_mel_features,
    }
