import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import orz_b200
from oracle import pyoracle
from tests import datagen
gen, n, level = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
data = getattr(datagen, gen)(n)
t = time.time()
try:
    got = orz_b200.encode(data, level)
    print("encode ok", time.time() - t, got == pyoracle.encode(data, level))
except Exception as e:
    print("encode failed:", e, time.time() - t)
