#!/usr/bin/env python3
"""Write a deck in this repo's canonical form: header comment kept, keys sorted by name, grouped
by prefix, values aligned.  usage: canon_deck.py deck.in [...]   (rewrites in place)"""
import sys


def canon(text):
    head, items = [], {}
    for line in text.splitlines():
        if line.lstrip().startswith('#') and not items:
            head.append(line.rstrip())
            continue
        body = line.split('#', 1)[0].strip()
        if not body or '=' not in body:
            continue
        k, v = body.split('=', 1)
        items[k.strip()] = ' '.join(v.split())
    width = max(len(k) for k in items)
    out = list(head)
    if not any('canonical form' in h for h in head):
        out += ['#', '# (canonical form: keys sorted by name; the deck reader does not depend on the order)']
    prev = None
    for k in sorted(items):
        pre = k.split('.', 1)[0]
        if pre != prev:
            out.append('')
            prev = pre
        out.append(f'{k.ljust(width)}  =  {items[k]}')
    return '\n'.join(out) + '\n'


if __name__ == '__main__':
    for path in sys.argv[1:]:
        text = canon(open(path).read())
        open(path, 'w').write(text)
