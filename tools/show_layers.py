import json, sys
d = json.load(open(sys.argv[1]))
tf = td = tw = 0
for r in d['rows']:
    tf += 4 * r['count'] * r['fprop_ms']; td += 2 * r['count'] * (r['dgrad_ms'] or 0); tw += 2 * r['count'] * r['wgrad_ms']
    print("%4d %4d %d %d %3d x%d | f %.3f (%4.0f TF) d %.3f (%4.0f) w %.3f (%4.0f)" % (
        r['cin'], r['cout'], r['k'], r['stride'], r['hin'], r['count'], r['fprop_ms'], r['fprop_tflops'],
        r['dgrad_ms'] or 0, r['dgrad_tflops'] or 0, r['wgrad_ms'], r['wgrad_tflops']))
print("per step: fprop %.1f dgrad %.1f wgrad %.1f total %.1f" % (tf, td, tw, tf + td + tw))
