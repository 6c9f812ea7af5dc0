// mlease_regression -- CLI of the host job layer:  mlease_regression [<job class>] <job config file>
// (default job class: Regression = Prepare -> AdmmTrain -> Test -> TestLoglik, jobs/Regression.java:37-98)
#include <cstdio>
#include <cstring>

#include "../../include/mlease_host.h"

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "[Usage]: mlease_regression [<job class>] <Job config path>\n"); return 64; }
  const char* job = argc >= 3 ? argv[1] : "Regression";
  const char* cfg = argc >= 3 ? argv[2] : argv[1];
  int rc = mlease_job_run(job, cfg);
  if (rc) fprintf(stderr, "%s failed: %s\n", job, mlease_job_last_error());
  return rc;
}
