import re, sys
src = open('include/xcp.h').read()
src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
out = {}
for m in re.finditer(r'(const char\*|long long|int)\s+(xcp_\w+)\s*\(([^)]*)\)\s*;', src):
    ret, name, args = m.groups()
    sig = ''
    args = args.strip()
    if args and args != 'void':
        for a in args.split(','):
            a = a.strip()
            if '*' in a: sig += 'p'
            elif a.startswith('long long'): sig += 'l'
            elif a.startswith('int'): sig += 'i'
            elif a.startswith('float'): sig += 'f'
            elif a.startswith('double'): sig += 'd'
            else: raise SystemExit('unknown arg '+a)
    out[name] = (ret, sig)
for k, (r, s) in out.items():
    print(f'    "{k}": "{s}",' + ('   # returns ' + r if r != 'int' else ''))
