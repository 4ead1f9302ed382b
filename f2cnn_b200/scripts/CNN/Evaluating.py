"""Drop-in for the feature front end of the reference's scripts/CNN/Evaluating.py.

EvaluateOneWavArray (reference :27-113) spends 25 s per 3 s utterance in a pure-Python dense
framing loop (:70-78) and ~1 s in per-frame normalizeInput (:79-80); here filterbank, envelope,
framing and normalisation are one GPU launch sequence (api.dense_frames).  Model loading,
prediction, the accuracy heuristic and plotting are Keras / matplotlib code of the reference
and are called exactly as the reference calls them (lazy imports; they need the reference tree
and its dependencies)."""
import glob
import os
import time
from configparser import ConfigParser
from shutil import copyfile

import numpy

from ...gammatone import filters
from ..processing.GammatoneFiltering import GetArrayFromWAV


def SNRdbToSNRlinear(SNRdb):
    """Reference :180-181 (a power ratio, applied to an amplitude by the caller -- kept)."""
    return 10 ** (SNRdb / 10.0)


def RMS(signal):
    """Reference :184-190.  numpy.square keeps the input dtype: on an int16 WAV array the
    squares overflow, exactly like the reference."""
    return numpy.sqrt(numpy.mean(numpy.square(signal)))


def PrepareInputFromArray(wavArray, framerate, config, LPF=False, CUTOFF=100, CENTER_FREQUENCIES=None,
                          FILTERBANK_COEFFICIENTS=None):
    """The hot part of EvaluateOneWavArray (:42-81): returns (input_data (nb, 2R+1, C) float64
    normalised per frame, CENTER_FREQUENCIES, FILTERBANK_COEFFICIENTS, STEP)."""
    from ... import api
    RADIUS = config.getint('CNN', 'RADIUS')
    SAMPPERIOD = config.getint('CNN', 'SAMPLING_PERIOD')
    USTOS = 1 / 1000000.
    if CENTER_FREQUENCIES is None:
        NCHANNELS = config.getint('FILTERBANK', 'NCHANNELS')
        lowcutoff = config.getint('FILTERBANK', 'LOW_FREQ')
        CENTER_FREQUENCIES = filters.centre_freqs(framerate, NCHANNELS, lowcutoff)
        FILTERBANK_COEFFICIENTS = filters.make_erb_filters(framerate, CENTER_FREQUENCIES)
    STEP = int(framerate * SAMPPERIOD * USTOS)
    print("Applying filterbank...")
    if not LPF:
        print("Extracting Envelope...")
    else:
        print("Extraction Envelope with {}Hz Low Pass Filter...".format(CUTOFF))
    print(LPF, CUTOFF)
    print("Generating input data for CNN...")
    input_data = api.dense_frames(wavArray, FILTERBANK_COEFFICIENTS, LPF, CUTOFF, RADIUS, STEP, normalize=True,
                                  dtype=numpy.float64)
    print("INPUT SHAPE:", input_data.shape)
    return input_data, CENTER_FREQUENCIES, FILTERBANK_COEFFICIENTS, STEP


def EvaluateOneWavArray(wavArray, framerate, wavFileName, model='last_trained_model', LPF=False, CUTOFF=100,
                        CENTER_FREQUENCIES=None, FILTERBANK_COEFFICIENTS=None):
    """Reference :27-113."""
    from ... import api
    from scripts.processing.LabelDataGenerator import ExtractLabel          # reference tree
    from scripts.processing.FBFileReader import ExtractFBFile               # reference tree
    from scripts.processing.PHNFileReader import ExtractPhonemes            # reference tree
    from scripts.plotting.PlottingCNN import PlotEnvelopesAndCNNResultsWithPhonemes  # reference tree
    config = ConfigParser()
    config.read('configF2CNN.conf')
    RADIUS = config.getint('CNN', 'RADIUS')
    NCHANNELS = config.getint('FILTERBANK', 'NCHANNELS')
    DOTSPERINPUT = RADIUS * 2 + 1

    labels = ExtractLabel(wavFileName, config)
    labels = [(entry[-4], entry[-1]) for entry in labels] if labels is not None else None

    input_data, CENTER_FREQUENCIES, FILTERBANK_COEFFICIENTS, STEP = PrepareInputFromArray(
        wavArray, framerate, config, LPF, CUTOFF, CENTER_FREQUENCIES, FILTERBANK_COEFFICIENTS)
    nb = input_data.shape[0]
    # the plots need the full-rate envelopes too (:108)
    envelopes = api.filterbank_envelope(wavArray, FILTERBANK_COEFFICIENTS, LPF, CUTOFF)

    print("Extracting Formants...")
    formants, sampPeriod = ExtractFBFile(os.path.splitext(wavFileName)[0] + '.FB')
    print("Extracting Phonemes...")
    phonemes = ExtractPhonemes(os.path.splitext(wavFileName)[0] + '.PHN')

    print("Evaluating the data with the pretrained model...")
    import keras
    model = keras.models.load_model(model)
    scores = model.predict(input_data.reshape(nb, DOTSPERINPUT, NCHANNELS, 1), verbose=1)
    simplified_scores = [1 if score[1] > score[0] else 0 for score in scores]
    keras.backend.clear_session()
    del model
    del input_data
    accuracy = None
    if labels is not None:
        accuracy = 0
        total_valid = 0
        for timepoint, score in enumerate(simplified_scores):
            for index in range(len(labels) - 1):
                before = labels[index][0]
                after = labels[index + 1][0]
                if before < timepoint < after and (abs(timepoint - before) < STEP or abs(timepoint - after) < STEP):
                    if abs(before - timepoint) <= abs(after - timepoint):
                        if score == labels[index][1]:
                            accuracy += 1
                    else:
                        if score == labels[index + 1][1]:
                            accuracy += 1
                    total_valid += 1
        accuracy /= total_valid
    print("Plotting...")
    PlotEnvelopesAndCNNResultsWithPhonemes(envelopes, scores, accuracy, CENTER_FREQUENCIES, phonemes, formants,
                                           wavFileName)


def EvaluateOneWavFile(file, LPF=False, CUTOFF=50, model='last_trained_model', CENTER_FREQUENCIES=None,
                       FILTERBANK_COEFFICIENTS=None):
    """Reference :116-137."""
    print('Using model', model)
    print("File:\t\t{}".format(file))
    framerate, wavArray = GetArrayFromWAV(file)
    EvaluateOneWavArray(wavArray=wavArray, framerate=framerate, LPF=LPF, CUTOFF=CUTOFF, wavFileName=file, model=model,
                        CENTER_FREQUENCIES=CENTER_FREQUENCIES, FILTERBANK_COEFFICIENTS=FILTERBANK_COEFFICIENTS)
    print("\t\t{}\tdone !".format(file))


def EvaluateRandom(count=None, LPF=False, CUTOFF=50):
    """Reference :140-177."""
    os.environ['TF_CPP_MIN_LOG_LEVEL'] = '3'
    TotalTime = time.time()
    if not os.path.isdir("graphs"):
        os.mkdir('graphs')
        os.mkdir(os.path.join('graphs', 'FallingOrRising'))
    wavFiles = glob.glob(os.path.join('resources', 'f2cnn', '*', '*.WAV'))
    print("\n###############################\nEvaluating network on {} WAV files in '{}'.".format(
        len(wavFiles), os.path.split(wavFiles[0])[0]))
    if not wavFiles:
        print("NO WAV FILES FOUND")
        exit(-1)
    config = ConfigParser()
    config.read('configF2CNN.conf')
    framerate = config.getint('FILTERBANK', 'FRAMERATE')
    nchannels = config.getint('FILTERBANK', 'NCHANNELS')
    lowcutoff = config.getint('FILTERBANK', 'LOW_FREQ')
    CENTER_FREQUENCIES = filters.centre_freqs(framerate, nchannels, lowcutoff)
    FILTERBANK_COEFFICIENTS = filters.make_erb_filters(framerate, CENTER_FREQUENCIES)
    if count is None:
        numpy.random.shuffle(wavFiles)
    elif count > 1:
        wavFiles = numpy.random.choice(wavFiles, count)
    for file in wavFiles:
        EvaluateOneWavFile(file, LPF=LPF, CUTOFF=CUTOFF, CENTER_FREQUENCIES=CENTER_FREQUENCIES,
                           FILTERBANK_COEFFICIENTS=FILTERBANK_COEFFICIENTS)
    print("Evaluating network on all files.")
    print('              Total time:', time.time() - TotalTime)
    print('')


def MixNoise(wavList, SNRdB):
    """Reference :199-200: noise = normal(scale=RMS(wav)/10^(dB/10)); output = noise + wav (float64)."""
    noise = numpy.random.normal(scale=RMS(wavList) / SNRdbToSNRlinear(SNRdB), size=wavList.shape[0])
    return noise + wavList


def EvaluateWithNoise(file, LPF=False, CUTOFF=100, model='last_trained_model', CENTER_FREQUENCIES=None,
                      FILTERBANK_COEFFICIENTS=None, SNRdB=-3):
    """Reference :193-221."""
    from scipy.io import wavfile
    print("File:\t\t{}".format(file))
    print("Appyling gaussian noise, new SNR is {SNR}dB".format(SNR=SNRdB))
    framerate, wavList = GetArrayFromWAV(file)
    output = MixNoise(wavList, SNRdB)
    os.makedirs(os.path.join('OutputWavFiles', 'addedNoise'), exist_ok=True)
    baseName = os.path.join('OutputWavFiles', 'addedNoise',
                            os.path.split(os.path.splitext(file)[0])[1]) + '{SNR}dB'.format(SNR=SNRdB)
    newPath = baseName + '.WAV'
    srcBasename = os.path.splitext(file)[0]
    wavfile.write(newPath, framerate, output)
    try:
        copyfile(srcBasename + '.FB', baseName + '.FB')
        copyfile(srcBasename + '.PHN', baseName + '.PHN')
        copyfile(srcBasename + '.WRD', baseName + '.WRD')
    except FileNotFoundError as e:
        print(e.strerror)
        print("No .FB or .PHN or .WRD files.")
    print('New noisy WAVE file saved as', newPath)
    EvaluateOneWavArray(output, framerate, newPath, model=model, LPF=LPF, CUTOFF=CUTOFF,
                        CENTER_FREQUENCIES=CENTER_FREQUENCIES, FILTERBANK_COEFFICIENTS=FILTERBANK_COEFFICIENTS)
    print("\t\t{}\tdone !".format(file))
