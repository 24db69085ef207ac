import json,sys
txt=open(sys.argv[1]).read()
dec=json.JSONDecoder(); i=0; objs=[]
while i < len(txt):
    j=txt.find('{',i)
    if j<0: break
    try:
        o,k=dec.raw_decode(txt[j:]); objs.append(o); i=j+k
    except Exception as e:
        i=j+1
for o in objs:
    print("==", o.get('case'), o.get('rows'), o.get('cols'), o.get('V'), o.get('box'))
    for k,v in o.items():
        if isinstance(v,dict) and 'mismatch' in v:
            if v['mismatch'] or '-v' in sys.argv:
                print("  %-45s mismatch %8d / %8d  max_rel %s  first %s" % (k, v['mismatch'], v['n'], v.get('max_rel'), v.get('first',[None])[0]))
        elif k not in ('case','rows','cols','V','box'): print("  ",k,v)
    print("  exact:", [k for k,v in o.items() if isinstance(v,dict) and v.get('mismatch')==0].__len__(), "of", [k for k,v in o.items() if isinstance(v,dict) and 'mismatch' in v].__len__())
