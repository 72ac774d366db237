import json, sys
b = json.loads(open(sys.argv[1]).read())
print('N', b['n_gpus'], 'value %.4g' % b['value'], 'ms/step %.3f host %.3f' % (b['ms_per_step'], b['ms_per_step_host']),
      'e2e %.4g %.3f ms' % (b['e2e']['value'], b['e2e']['ms_per_step']), 'e2e_op %.4g' % b['e2e_operator']['value'], b['clocks'])
print('by rank', b.get('ms_per_step_by_rank'))
