"""Constants of the reference's pylamp_const.py (values are part of the contract).

Restated from /root/reference/pylamp_const.py:3-46: axis indices, gravity, time
constants, gas constant, the 13 tracer-column indices and the absolute fence EPS.
"""
DEBUG = 3
DIM = 2
IZ = 0
IX = 1
IY = 2
IP = DIM

G = [9.81, 0]
SECINYR = 60 * 60 * 24 * 365.25
SECINKYR = SECINYR * 1e3
SECINMYR = SECINYR * 1e6
GASR = 8.31446

NFTRAC = 13
TR_RHO = 0
TR_ETA = 1
TR_MRK = 2
TR_TMP = 3
TR_HCD = 4
TR_HCP = 5
TR_RH0 = 6
TR_ALP = 7
TR_MAT = 8
TR_ACE = 9
TR_ET0 = 10
TR_IHT = 11
TR__ID = 12

EPS = 2 ** (-10)
