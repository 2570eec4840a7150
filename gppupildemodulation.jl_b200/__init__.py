"""gppd-b200: B200-native demodulateall hot path of GPPupilDemodulation.jl."""
