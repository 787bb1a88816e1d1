// oracle/oracle_realtime.cpp — TEST INFRASTRUCTURE ONLY (see oracle.cpp header).
// Placeholder translation unit for the realtime integrator restatement (pt_raygen_realtime.rgen).
