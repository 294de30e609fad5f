from .agh.problem_agh import AGH
