#!/usr/bin/env python
"""Generates the coefficient tables of Fukushima's (2015, Appl. Math. Comput. 259, 708) piecewise
minimax rational approximations of the Fermi-Dirac integrals F_{-3/2}, F_{-1/2}, F_{1/2}.

The constants are DATA of a published algorithm; XNet carries them inline in nested Horner form
(source/xnet_fd.F90, functions fdm3h / fdm1h / fd1h).  This script reads those expressions, turns every
interval into two plain coefficient arrays (numerator, denominator; ascending powers, exactly the
Horner nesting of the published code) and writes the same table twice:

    oracle/ox_fd_tables.h            (C, test infrastructure)
    xnet_b200/csrc/xb_fd_tables.cuh  (CUDA __constant__, product)

The evaluators that use the tables are hand written (oracle/ox_eos.c, xnet_b200/csrc/xb_eos.cuh).
Run in the development container only (needs /root/reference); the outputs are committed.
"""
import ast
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else '/root/reference/source/xnet_fd.F90'
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')


def function_body(src, name):
    m = re.search(r'[Ff]unction %s\([xy]\)(.*?)End Function %s' % (name, name), src, re.S)
    return m.group(1)


def branches(body):
    """-> list of (condition text or None, {var: expr}, fd expr) in source order"""
    body = re.sub(r'&\s*\n', ' ', body)
    out, cond, defs = [], None, {}
    for line in body.splitlines():
        line = line.strip()
        m = re.match(r'(?:Else)?If\((.*?)\)\s*Then', line)
        if m:
            cond, defs = m.group(1), {}
            continue
        if line == 'Else':
            cond, defs = None, {}
            continue
        m = re.match(r'(ex|t|s|w|fd)=(.*)', line)
        if m:
            expr = re.sub(r'([\d.])d([+-]?\d+)', r'\1e\2', m.group(2))
            if m.group(1) == 'fd':
                out.append((cond, dict(defs), expr))
            else:
                defs[m.group(1)] = expr
    return out


def num(node):
    if isinstance(node, ast.Constant):
        return float(node.value)
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
        return -num(node.operand)
    raise ValueError(ast.dump(node))


def horner(node):
    """c0 + v*(c1 + v*(...)) -> (v, [c0, c1, ...]); a bare variable is the monic tail [0, 1]"""
    if isinstance(node, ast.Name):
        return node.id, [0.0, 1.0]
    if isinstance(node, ast.BinOp) and isinstance(node.op, (ast.Add, ast.Sub)):
        sign = 1.0 if isinstance(node.op, ast.Add) else -1.0
        c0 = num(node.left)
        r = node.right
        if isinstance(r, ast.Name):
            return r.id, [c0, sign]
        assert isinstance(r, ast.BinOp) and isinstance(r.op, ast.Mult) and isinstance(r.left, ast.Name), ast.dump(r)
        v = r.left.id
        try:
            return v, [c0, sign * num(r.right)]
        except ValueError:
            v2, cs = horner(r.right)
            assert v2 == v
            return v, [c0] + [sign * c for c in cs]
    return None, [num(node)]


def find_rational(node):
    """first Div whose both sides are Horner forms, searched depth first -> (P, Q, vP, vQ) or None"""
    for n in ast.walk(node):
        if isinstance(n, ast.BinOp) and isinstance(n.op, ast.Div):
            try:
                left = n.left
                if isinstance(left, ast.UnaryOp) and isinstance(left.op, ast.USub):
                    left = left.operand
                # (w*P)/Q or (ex*P)/Q: strip the leading factor
                if isinstance(left, ast.BinOp) and isinstance(left.op, ast.Mult) and isinstance(left.left, ast.Name):
                    left = left.right
                vp, p = horner(left)
                vq, q = horner(n.right)
                if len(p) > 2 and len(q) > 2:
                    return p, q, vp, vq
            except (ValueError, AssertionError):
                continue
    return None


def tables(name, src):
    rows = []
    for cond, defs, expr in branches(function_body(src, name)):
        if cond and 'y>' in cond:
            continue   # reflection formula of the integer orders: hand written in the evaluators
        tree = ast.parse(expr, mode='eval').body
        rat = find_rational(tree)
        lead = re.match(r'\s*-\(', expr) is not None
        if rat is None:
            # asymptotic series without denominator (fdm1h, x >= 40): sqrt(x)*factor*(1 - w*P(t))
            inner = [n for n in ast.walk(tree) if isinstance(n, ast.BinOp) and isinstance(n.op, ast.Mult)
                     and isinstance(n.left, ast.Name) and n.left.id == 'w']
            vp, p = horner(inner[0].right)
            rat = (p, [1.0], vp, None)
        p, q, vp, vq = rat
        # additive constant of the x < -2 branch: ex*(c0 +- ex*P/Q)
        c0 = 0.0
        sign = -1.0 if lead else 1.0
        if cond and '-2' in cond:
            m = re.match(r'ex\*\(([-0-9.e+]+)\s*([+-])ex\*', expr.replace(' ', ''))
            c0 = float(m.group(1)); sign = 1.0 if m.group(2) == '+' else -1.0
        elif cond is None:
            m = re.search(r'\(1\.e0\s*([+-])w\*', expr)
            sign = 1.0 if m.group(1) == '+' else -1.0
        rows.append(dict(cond=cond, defs=defs, p=p, q=q, c0=c0, sign=sign, expr=expr))
    return rows


def emit(rows_by_fn, qualifier, guard):
    out = ['// GENERATED by tools/gen_fd_tables.py -- coefficient tables of Fukushima (2015) for the Fermi-Dirac',
           '// integrals of order -3/2, -1/2, 1/2 (XNet: source/xnet_fd.F90).  Interval k of function f:',
           '//   numerator   P[k][0..np-1], denominator Q[k][0..nq-1], ascending powers (Horner from the top);',
           '//   c0/sign     additive constant and sign of the rational part where the closed form has them.',
           '// Intervals: 0: x<-2, 1: x<0, 2: x<2, 3: x<5, 4: x<10, 5: x<20, 6: x<40, 7: x>=40.  The integer orders',
           '// (fd0h, fd2h, ... = k = 0, 1, 2, 3, 4) have the first two intervals only, on x = -|y|, plus a reflection formula.',
           '#pragma once', '#define %s 1' % guard, '#define FD_MAXC 10']
    for fn, rows in rows_by_fn.items():
        n_iv = len(rows)
        assert n_iv in (2, 8), (fn, n_iv)
        for key in ('p', 'q'):
            out.append('%s double %s_%s[%d][FD_MAXC] = {' % (qualifier, fn, key, n_iv))
            for r in rows:
                cs = r[key] + [0.0] * (10 - len(r[key]))
                assert len(cs) == 10, (fn, len(r[key]))
                out.append('  {' + ', '.join(repr(c) for c in cs) + '},')
            out.append('};')
            out.append('%s int %s_n%s[%d] = {%s};' % (qualifier, fn, key, n_iv, ', '.join(str(len(r[key])) for r in rows)))
        out.append('%s double %s_c0 = %r;' % (qualifier, fn, rows[0]['c0']))
        out.append('%s double %s_sign[%d] = {%s};' % (qualifier, fn, n_iv, ', '.join(repr(r['sign']) for r in rows)))
    return '\n'.join(out) + '\n'


def main():
    src = open(REF).read()
    rows = {fn: tables(fn, src) for fn in ('fdm3h', 'fdm1h', 'fd1h', 'fd0h', 'fd2h', 'fd4h', 'fd6h', 'fd8h')}
    for fn, rs in rows.items():
        for k, r in enumerate(rs):
            print(fn, k, r['cond'], r['defs'], 'np', len(r['p']), 'nq', len(r['q']), 'c0', r['c0'], 'sign', r['sign'])
    open(os.path.join(ROOT, 'oracle', 'ox_fd_tables.h'), 'w').write(emit(rows, 'static const', 'OX_FD_TABLES'))
    open(os.path.join(ROOT, 'xnet_b200', 'csrc', 'xb_fd_tables.cuh'), 'w').write(
        emit(rows, '__device__ const', 'XB_FD_TABLES'))


if __name__ == '__main__':
    main()
